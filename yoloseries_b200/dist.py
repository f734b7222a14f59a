"""Image-sharded post-processing across the GPUs of one box + the one collective the path has.

Images are independent units (trainer/eval_yolov5.py:268 loops per image), so rank r simply owns images
[r*b/W, (r+1)*b/W) and no data-path collective is needed until the end, where the kept detections of every rank are
all-gathered for mAP evaluation.  The reference never gathers detections (its val sampler is not even rank-sliced,
dataset/data_sampler.py:185-192); the closest thing it has is the unused pickle-over-Gloo ``all_gather`` of
utils/dist.py:176-211.  Here the message is a fixed-stride buffer -- rows padded to max_det plus the per-image counts --
so there is no size pre-exchange, and on a B200 box it travels over NVLink/NVSwitch through NCCL.
Works with any torch.distributed backend (``gloo`` on CPU tensors is what the CPU tests use).
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """[lo, hi) of the images rank ``rank`` owns; the first ``batch % world`` ranks take one extra image."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_detections(dets, cnt, per_rank):
    """(b_local, max_det, 6) f32 + (b_local,) i32 -> one flat f32 buffer sized for ``per_rank`` images."""
    b, max_det, _ = dets.shape
    buf = torch.zeros(per_rank * (max_det * 6 + 1), dtype=torch.float32, device=dets.device)
    buf[: b * max_det * 6] = dets.reshape(-1)
    cnt_f = torch.full((per_rank,), -2.0, dtype=torch.float32, device=dets.device)  # -2 marks padding images
    cnt_f[:b] = cnt.to(torch.float32)
    buf[per_rank * max_det * 6:] = cnt_f
    return buf


def gather_detections(dets, cnt, batch, group=None):
    """All-gather the kept detections of an image-sharded batch.

    ``dets`` (b_local, max_det, 6) and ``cnt`` (b_local,) are this rank's shard (in shard_bounds order); ``batch`` is the
    global image count.  Returns (dets_all (batch, max_det, 6), cnt_all (batch,)) on every rank; ``cnt == -1`` keeps
    meaning "the reference returns None for this image".
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return dets, cnt
    max_det = dets.shape[1]
    per_rank = (batch + world - 1) // world
    send = pack_detections(dets, cnt, per_rank)
    recv = torch.empty(world * send.numel(), dtype=torch.float32, device=dets.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, -1)
    rows = recv[:, : per_rank * max_det * 6].reshape(world, per_rank, max_det, 6)
    cnts = recv[:, per_rank * max_det * 6:].to(torch.int32)
    out_d, out_c = [], []
    for r in range(world):
        lo, hi = shard_bounds(batch, r, world)
        out_d.append(rows[r, : hi - lo])
        out_c.append(cnts[r, : hi - lo])
    return torch.cat(out_d, 0), torch.cat(out_c, 0)


# ----------------------------------------------------------------------------------------------------------------------
# The gather without a collective launch: symmetric receive buffers mapped over NVLink (CUDA IPC), written by the NMS
# kernel itself (include/ysb_postproc.h "multi-GPU", csrc/gather_kernels.cu).
# ----------------------------------------------------------------------------------------------------------------------
import ctypes
import os

from . import _lib


def exchange_bytes(payload: bytes, group=None):
    """All-gather one small bytes object per rank (IPC handles) through torch.distributed; works on gloo and NCCL."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [payload]
    out = [None] * world
    dist.all_gather_object(out, payload, group=group)
    return out


def bind_to_gpu_numa(local_rank):
    """Pin the calling process to the CPU cores next to GPU ``local_rank`` (NVML's ideal affinity) so that pinned host
    buffers allocated afterwards are first-touched on that GPU's NUMA node: with every rank's staging memory on node 0,
    H2D throughput stops scaling beyond two GPUs.  Best effort; returns the core count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        return None
    return None


def default_lanes(batch, world):
    """Batches in flight per rank (one CUDA stream each) when the caller does not say.

    Large batches: 4 hide the NMS kernel behind the filter kernels of the next batches; with peers a lane also stays busy
    until every rank's rows have landed, which 6 lanes cover.  Batches below 64 images leave SMs idle per launch and want 8.
    Measured on B200 (YOLOv5s 640, images/s): b=64 4 / 6 lanes 738 k / 722 k on one GPU, 0.0944 / 0.0891 ms per step on
    two; b=32 603 k / 694 k (4 / 6); b=16 550 k / 618 k / 654 k (4 / 6 / 8); YOLOv5x-1280 b=16 174.5 k / 185.2 k / 187.7 k;
    b=8 476 k / 480 k / 485 k."""
    if int(batch) < 64:
        return 8
    return 4 if int(world) == 1 else 6


class _DevicePointer:
    """``__cuda_array_interface__`` holder: a zero-copy torch view of memory the C ABI allocated."""

    def __init__(self, ptr, shape, typestr, strides, owner):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr,
                                         "strides": tuple(strides), "version": 3}
        self._owner = owner


class DetectionGather:
    """Symmetric receive buffers of one rank + the mappings of every peer's buffer.

    ``mode``: 'p2p' (the NMS kernel stores rows into every peer's slot), 'local' (world == 1), decided at construction;
    a failed IPC mapping raises (callers that want the NCCL path ask for it explicitly, see ShardedPostProcessor)."""

    def __init__(self, batch, max_det, slots, device, group=None):
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > _lib.YSB_MAX_PEERS:
            raise ValueError(f"at most {_lib.YSB_MAX_PEERS} ranks")
        self.batch, self.max_det, self.slots, self.device = int(batch), int(max_det), int(slots), torch.device(device)
        nbytes = ctypes.c_size_t()
        _lib.check(self.lib.ysb_gather_buffer_bytes(self.world, self.slots, self.batch, self.max_det, ctypes.byref(nbytes)),
                   "ysb_gather_buffer_bytes")
        self.nbytes = nbytes.value
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(_lib.YSB_IPC_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ysb_gather_alloc(self.nbytes, ctypes.byref(own), handle), "ysb_gather_alloc")
        self._own = own.value
        self._opened = []
        g = _lib.YsbGather()
        g.world, g.rank, g.slots, g.batch, g.max_det = self.world, self.rank, self.slots, self.batch, self.max_det
        g.d_buf[self.rank] = self._own
        handles = exchange_bytes(handle.raw, group)
        with torch.cuda.device(self.device):
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                p = ctypes.c_void_p()
                _lib.check(self.lib.ysb_gather_open(h, ctypes.byref(p)), f"ysb_gather_open(rank {r})")
                self._opened.append(p.value)
                g.d_buf[r] = p.value
        self.g = g
        rs, cs = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.ysb_gather_strides(ctypes.byref(g), ctypes.byref(rs), ctypes.byref(cs)), "ysb_gather_strides")
        self._views = []
        for s in range(self.slots):
            rp, cp = ctypes.c_void_p(), ctypes.c_void_p()
            _lib.check(self.lib.ysb_gather_slot_views(ctypes.byref(g), s, ctypes.byref(rp), ctypes.byref(cp)),
                       "ysb_gather_slot_views")
            with torch.cuda.device(self.device):
                rows = torch.as_tensor(_DevicePointer(rp.value, (self.world, self.batch, self.max_det, 6), "<f4",
                                                      (rs.value, self.max_det * 24, 24, 4), self), device=self.device)
                cnt = torch.as_tensor(_DevicePointer(cp.value, (self.world, self.batch), "<i4", (cs.value, 4), self),
                                      device=self.device)
            self._views.append((rows, cnt))
        if self.world > 1:
            dist.barrier(group)   # every rank has mapped every buffer before anybody's kernels write into them

    def views(self, slot):
        """(rows (world, batch, max_det, 6) f32, counts (world, batch) i32) of this rank's receive slot."""
        return self._views[slot]

    def error(self):
        e = ctypes.c_uint32()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ysb_gather_error(ctypes.byref(self.g), ctypes.byref(e)), "ysb_gather_error")
        return e.value

    def close(self):
        """Collective: every rank stops using the buffers, then unmaps and frees."""
        if self._own is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if self.world > 1 and dist.is_initialized():
                dist.barrier(self.group)
            for p in self._opened:
                self.lib.ysb_gather_close(p)
            self._views = []
            self.lib.ysb_gather_free(self._own)
        self._own, self._opened = None, []


class ShardedPostProcessor:
    """The multi-GPU path as an API: every rank post-processes its own image shard, ``lanes`` batches in flight (default:
    4, 6 with peers, 8 for batches below 64 images), and the
    kept detections of ALL ranks end up on EVERY rank.

        spp = ShardedPostProcessor("yolov5", hyp, batch=8, img_h=640, img_w=640, anchors=anchors)
        t = spp.submit(heads)                  # filter -> NMS (+ stores into every peer's slot) -> arrival wait, async
        rows, cnt = spp.gathered(t)            # (world, batch, max_det, 6), (world, batch) device views, stream-ordered
        dets = spp.result(t)                   # list over world*batch images: CPU (K, 6) tensors / None

    Batch i runs on CUDA stream i % lanes with its own key/count buffers, so the latency-bound NMS kernel of one batch
    overlaps the HBM-bound filter kernels of the next ones.  gather = 'p2p': no collective launch at all (the NMS kernel
    writes over NVLink, flow control by two tiny kernels); 'nccl': one all_gather_into_tensor per batch on the lane's
    stream (fallback when CUDA IPC is unavailable); world == 1: plain local buffers.  With ``graph`` (default: batch <=
    16) the per-lane launch sequence is captured once per set of head tensors and replayed.
    A ticket's buffers are reused ``lanes`` submissions later.
    """

    def __init__(self, family, hyp, batch, img_h, img_w, anchors=None, lanes=None, group=None, gather="auto", graph=None,
                 compute_metric=False, device=None):
        from .engine import PostProcessor
        self.family, self.batch, self.img_h, self.img_w = family, int(batch), int(img_h), int(img_w)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lanes = max(1, int(lanes)) if lanes else default_lanes(self.batch, self.world)
        self.pp = PostProcessor(family, hyp, anchors=anchors, compute_metric=compute_metric)
        self.lib = _lib.load()
        self.max_det = int(hyp["max_predictions_per_img"])
        self.graph = (self.batch <= 16) if graph is None else bool(graph)
        want = os.environ.get("YSB_GATHER", gather)
        self.mode = "local" if self.world == 1 else ("p2p" if want in ("auto", "p2p") else "nccl")
        self.dg = None
        if self.mode == "p2p":
            ok = torch.ones(1, device=self.device)
            try:
                self.dg = DetectionGather(self.batch, self.max_det, self.lanes, self.device, group)
            except Exception as e:   # every rank must take the same path: agree on it below
                self._p2p_error = repr(e)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if ok.item() == 0:
                if want == "p2p":
                    raise RuntimeError("P2P detection gather unavailable: " + getattr(self, "_p2p_error", "a peer failed"))
                if self.dg is not None:
                    self.dg.close()
                    self.dg = None
                self.mode = "nccl"
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)]
        self._ent = None
        self._slots = None
        self._graphs = {}
        self._warmed = False
        self._next = 0

    # ---- buffers ---------------------------------------------------------------------------------------------------
    def _prepare(self, flat):
        if self._ent is not None:
            return
        self._ent = self.pp._prepare(flat, self.batch, self.img_h, self.img_w, _lib.INPUT_RAW_HEADS)
        n = self._ent["N"] * (int(self.pp.hyp["num_class"]) if self.pp.hyp["mutil_label"] else 1)
        self._key_cap = n
        b, d, dev = self.batch, self.max_det, self.device
        self._slots = []
        cur = torch.cuda.current_stream(dev)
        for s in range(self.lanes):
            # A lane's buffers are allocated UNDER the lane's stream: the caching allocator hands out blocks whose last
            # use may still be pending on the stream they were freed on, which is only safe for work enqueued on that
            # same stream.  (Allocated on the caller's stream and written from the lane, a key buffer could share memory
            # with a temporary of a still-running producer kernel: corrupted keys -> wild candidate indices in the NMS
            # kernel.)  The lane also waits for everything the caller has enqueued so far.
            self.streams[s].wait_stream(cur)
            with torch.cuda.stream(self.streams[s]):
                sl = {"keys": torch.empty((b, n), dtype=torch.int64, device=dev),
                      "counts": torch.zeros((b, 4), dtype=torch.int32, device=dev),
                      "idx": torch.empty((b, d), dtype=torch.int32, device=dev),
                      "done": torch.cuda.Event()}
                if self.mode == "p2p":
                    sl["rows"], sl["cnt"] = self.dg.views(s)
                elif self.mode == "nccl":
                    sl["send"] = torch.zeros(b * d * 6 + b, dtype=torch.float32, device=dev)
                    sl["recv"] = torch.empty((self.world, b * d * 6 + b), dtype=torch.float32, device=dev)
                    sl["rows"] = sl["recv"][:, : b * d * 6].view(self.world, b, d, 6)
                    sl["cnt"] = sl["recv"][:, b * d * 6:].view(torch.int32)
                else:
                    sl["rows"] = torch.zeros((1, b, d, 6), dtype=torch.float32, device=dev)
                    sl["cnt"] = torch.zeros((1, b), dtype=torch.int32, device=dev)
            self._slots.append(sl)

    def _enqueue(self, ptrs, nheads, lane, st, events=None):
        """The launch sequence of one batch on stream ``st`` (plain C-ABI calls: capturable into a CUDA graph)."""
        lib, sl, params = self.lib, self._slots[lane], self._ent["params"]
        sp = ctypes.c_void_p(st.cuda_stream)
        if self.mode == "p2p":
            # first thing of the step: tell the peers this slot's previous contents have been consumed (they check it
            # right before storing into it, a filter + NMS pass from now)
            g = ctypes.byref(self.dg.g)
            _lib.check(lib.ysb_gather_begin(g, lane, None, 0, sp), "ysb_gather_begin")
        if events:
            events[0].record(st)
        _lib.check(lib.ysb_filter_candidates(ctypes.byref(params), ptrs, nheads, sl["keys"].data_ptr(), self._key_cap,
                                             sl["counts"].data_ptr(), sp), "ysb_filter_candidates")
        if events:
            events[1].record(st)
        if self.mode == "p2p":
            _lib.check(lib.ysb_select_nms_gather(ctypes.byref(params), ptrs, nheads, sl["keys"].data_ptr(), self._key_cap,
                                                 sl["counts"].data_ptr(), g, lane, sl["idx"].data_ptr(), sp),
                       "ysb_select_nms_gather")
            if events:
                events[2].record(st)
            _lib.check(lib.ysb_gather_wait(g, lane, sp), "ysb_gather_wait")
        else:
            if self.mode == "nccl":
                rows_ptr, cnt_ptr = sl["send"].data_ptr(), sl["send"].data_ptr() + self.batch * self.max_det * 24
            else:
                rows_ptr, cnt_ptr = sl["rows"].data_ptr(), sl["cnt"].data_ptr()
            _lib.check(lib.ysb_select_nms(ctypes.byref(params), ptrs, nheads, sl["keys"].data_ptr(), self._key_cap,
                                          sl["counts"].data_ptr(), rows_ptr, sl["idx"].data_ptr(), cnt_ptr, sp),
                       "ysb_select_nms")
            if events:
                events[2].record(st)

    # ---- public ----------------------------------------------------------------------------------------------------
    def submit(self, heads, events=None, sync_input=True):
        """Enqueue one batch.  ``events``: optional 3 timing events (start, filter done, NMS done) recorded on the lane's
        stream.  ``sync_input=False`` skips the wait on the caller's current stream (heads known to be ready)."""
        from .engine import flatten_heads
        flat = flatten_heads(self.family, heads)
        if flat[0].shape[0] != self.batch:
            raise ValueError(f"expected {self.batch} images per rank, got {flat[0].shape[0]}")
        self._prepare(flat)
        lane = self._next % self.lanes
        self._next += 1
        st = self.streams[lane]
        if sync_input:
            st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.device(self.device):
            key = (lane,) + tuple(t.data_ptr() for t in flat)
            if self.graph and events is None and not self._warmed:
                # the very first batch runs as plain launches (lazy module load, function attributes) and counts as a
                # normal step; graphs are captured from the second submission on
                self._enqueue(_lib.head_pointer_array(flat), len(flat), lane, st, None)
                torch.cuda.synchronize(self.device)
                self._warmed = True
            elif self.graph and events is None:
                gr = self._graphs.get(key)
                if gr is None:
                    gr = self._capture(flat, lane, st)
                    self._graphs[key] = gr
                with torch.cuda.stream(st):
                    gr.replay()
            else:
                self._enqueue(_lib.head_pointer_array(flat), len(flat), lane, st, events)
            if self.mode == "nccl":
                with torch.cuda.stream(st):
                    sl = self._slots[lane]
                    dist.all_gather_into_tensor(sl["recv"].view(-1), sl["send"], group=self.group)
        self._slots[lane]["done"].record(st)
        if sync_input:
            for t in flat:
                t.record_stream(st)
        return lane

    def _capture(self, flat, lane, st):
        ptrs = _lib.head_pointer_array(flat)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
            self._enqueue(ptrs, len(flat), lane, st)
        g._ysb_keep = (ptrs, flat)
        return g

    def gathered(self, ticket):
        """Device views (rows (world, batch, max_det, 6), counts (world, batch)); the caller's current stream is made to
        wait for the batch (and, p2p, for every peer's rows)."""
        sl = self._slots[ticket]
        torch.cuda.current_stream(self.device).wait_event(sl["done"])
        return sl["rows"], sl["cnt"]

    def result(self, ticket):
        """list over the world*batch images in rank order: CPU float32 (K, 6) tensors, or None (a4.3 contract)."""
        rows, cnt = self.gathered(ticket)
        cnt_h = cnt.cpu()
        kmax = max(int(cnt_h.max().item()), 1)
        rows_h = rows[:, :, :kmax].cpu()
        out = []
        for r in range(rows_h.shape[0]):
            for i in range(self.batch):
                c = int(cnt_h[r, i])
                out.append(None if c < 0 else rows_h[r, i, :c].clone())
        return out

    def drain(self):
        """The caller's current stream waits for everything submitted so far."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)

    def check(self):
        """Raises if a flow-control spin of the p2p gather timed out (host sync)."""
        if self.dg is not None:
            e = self.dg.error()
            if e:
                raise RuntimeError(f"detection gather timed out waiting for a peer (code {e})")

    def close(self):
        """Waits for the lanes, then releases the buffers (collective when the p2p gather is in use)."""
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
        self._graphs = {}
        if self.dg is not None:
            self.dg.close()
            self.dg = None
        self._slots, self._ent = None, None
