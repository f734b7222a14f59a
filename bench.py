#!/usr/bin/env python
"""bench.py -- post-processing throughput (images/s) of the B200 engine on BASELINE.json's headline config.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU algorithm on the host cores)

A "step" is one pass of the hot path (filter/compaction kernel + select/sort/NMS kernel) over one batch of
synthetic YOLOv5s 640x640 head tensors (3 levels, 25,200 candidates/image, 80 classes, conf=cls=0.001, iou=0.65,
max_det=300, class-aware).  Every rank owns `--batch` images (weak scaling: BASELINE config[1]'s 64 images per
GPU); at N > 1 the step ends with the all-gather of the padded kept detections (the only collective on the path).

Output: ONE JSON line on rank 0 (see the keys in main()).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "post-proc images/s (YOLOv5s 640^2, conf=0.001)"
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture
NCU_TRAFFIC_BYTES = {("yolov5", 640, 64, "dense"): 525095552 + 8164096}  # mean of the two captured launches
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per rank per step")
    ap.add_argument("--family", default="yolov5")
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--dist", default="dense", choices=["dense", "sparse", "crowd"])
    ap.add_argument("--pipeline", type=int, default=3, help="0: serial; 1: NMS kernel of batch i overlaps the filter kernel of batch i+1 on a side stream; 2: same, side stream at high priority; 3: two independent lanes (step i entirely on stream i %% 2)")
    ap.add_argument("--lanes", type=int, default=4, help="streams (= batches in flight) of --pipeline 3")
    ap.add_argument("--graph", type=int, default=0, help="replay one captured CUDA graph (memset + filter + NMS) per step and lane")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-images", type=int, default=2, help="images in the bounded CPU-baseline sample")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": f"{args.family} {args.img}x{args.img} decode+filter+top-k+class-aware NMS, "
                    f"{args.batch} images/GPU/step, 80 classes, conf=cls=0.001, iou=0.65, max_det=300, "
                    f"postprocess_bbox=true, distribution={args.dist}",
        "images_per_gpu_per_step": args.batch,
        "global_batch": args.batch * world,
        "distribution": args.dist,
        "parallelism": f"image-sharded x{world}" + (", all-gather of kept detections" if world > 1 else ""),
        "l2": "inputs per step (548 MB at 64 images) exceed the 126 MB L2; streamed once per step",
    }


def bench_hyp(family="yolov5"):
    from yoloseries_b200 import synth
    hyp = synth.map_profile_hyp(num_class=80)
    if family == "fcos":  # config/train_fcos.yaml:110-116
        hyp.update(cls_threshold=0.2, iou_threshold=0.35, max_predictions_per_img=100)
    return hyp


# ----------------------------------------------------------------------------------------------------------------
# clocks: sampled with NVML while the timed workload is running
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._armed = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            if self._armed.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = get_reasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def arm(self, on):
        (self._armed.set if on else self._armed.clear)()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1)
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ----------------------------------------------------------------------------------------------------------------
# CPU legs (oracle = plain C/numpy restatement of the reference's algorithm; the only place bench.py touches it)
# ----------------------------------------------------------------------------------------------------------------
def cpu_heads(n_images, args, seed=4321):
    from yoloseries_b200 import synth
    return [h.numpy() for h in synth.make_heads(args.family, n_images, args.img, args.img, 80, args.dist, seed, "cpu")]


def cpu_images_per_second(heads_np, threads):
    """Time the reference's CPU algorithm (decode -> filter -> FULL greedy NMS loop, no early stop -> post-filter)
    on the given synthetic images of the bench workload using `threads` host threads (one image per task)."""
    from concurrent.futures import ThreadPoolExecutor

    import oracle

    hyp = bench_hyp()
    n_images = heads_np[0].shape[0]

    def one(i):
        dec = oracle.decode_yolov5([h[i:i + 1] for h in heads_np])
        return oracle.evaluator_nms("yolov5", dec, hyp, full_nms=True)

    t0 = time.perf_counter()
    if threads <= 1:
        for i in range(n_images):
            one(i)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, range(n_images)))
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def parity_counters(args, pp, heads, slot, hyp, n_check=2):
    """North-star parity report on the first images of the bench batch (oracle = checker only, outside the timed
    region): decoded max relative error vs the numpy restatement, kept rows / candidate indices vs the oracle applied
    to the GPU-decoded tensor, and how many visited IoUs sit within 1e-6 of the threshold."""
    import numpy as np
    import oracle
    import torch

    sub = [h[:n_check].contiguous() for h in heads]
    dec_gpu = pp.decode(sub, args.img, args.img).cpu().numpy()
    dec_ref = oracle.decode_yolov5([h.cpu().numpy() for h in sub])
    rel = np.abs(dec_gpu - dec_ref) / np.maximum(np.abs(dec_ref), 1.0)
    want = oracle.evaluator_nms("yolov5", dec_gpu, hyp)
    torch.cuda.synchronize()
    rows_ok = idx_ok = True
    near = 0
    cnt = slot.cnt[:n_check].cpu().numpy()
    for i, w in enumerate(want):
        k = int(cnt[i])
        got_rows = slot.dets[i, :max(k, 0)].cpu().numpy()
        got_idx = slot.idx[i, :max(k, 0)].cpu().numpy()
        if w.rows is None:
            rows_ok &= k < 0
            continue
        rows_ok &= got_rows.shape == w.rows.shape and bool(np.array_equal(got_rows, w.rows))
        idx_ok &= bool(np.array_equal(got_idx, w.cand_index))
        order = np.lexsort((np.arange(len(w.nms_scores)), -w.nms_scores.astype(np.float64)))[:4096]
        iou = oracle.numba_iou(w.nms_boxes[np.asarray(w.keep, dtype=np.int64)], w.nms_boxes[order])
        near += int(np.sum(np.abs(iou - hyp["iou_threshold"]) < 1e-6))
    return {"images_checked": n_check, "decoded_max_rel_err": float(rel.max()), "decoded_within_1e-5": bool(rel.max() <= 1e-5),
            "kept_rows_bit_exact": bool(rows_ok), "kept_indices_bit_exact": bool(idx_ok),
            "iou_within_1e-6_of_threshold": near}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if args.family != "yolov5":
        print(json.dumps({"impl": "reference", "unavailable": "reference arm implemented for the yolov5 headline config"}))
        return
    import oracle
    oracle.load_library()
    threads = os.cpu_count() or 1
    budget_s = 240.0
    warm = min(args.warmup, 1)
    t_first = None
    heads_np = cpu_heads(threads, args)  # generated once; every step processes the same `threads` images
    for _ in range(warm):
        _, t_first = cpu_images_per_second(heads_np, threads)
    done, total_t = 0, 0.0
    for k in range(args.steps):
        _, dt = cpu_images_per_second(heads_np, threads)
        done += 1
        total_t += dt
        if total_t + (t_first or 0.0) > budget_s:
            break
    value = threads * done / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": done,
        "steps_requested": args.steps, "warmup": warm, "ms_per_step": 1e3 * total_t / done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (+f64 IoU)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{threads} images per step, one per host thread, full greedy NMS loop as "
                                   "utils/nms.py runs it (no early stop); C/numpy port of the reference algorithm "
                                   "(the reference itself is Python+numba and cannot travel to this box)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes

    import torch
    import torch.distributed as dist

    from yoloseries_b200 import _lib, synth
    from yoloseries_b200.engine import PostProcessor, flatten_heads
    if not os.path.exists(_lib.LIB_PATH):  # fresh checkout: the built library is git-ignored
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            from yoloseries_b200 import build as ysb_build
            print("bench.py: libysb_postproc.so missing, building it with nvcc ...", file=sys.stderr)
            ysb_build.build()
        else:  # the other ranks wait for rank 0's build (the link is renamed into place atomically)
            deadline = time.time() + 900
            while not os.path.exists(_lib.LIB_PATH) and time.time() < deadline:
                time.sleep(1.0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    hyp = bench_hyp(args.family)
    heads = synth.make_heads(args.family, args.batch, args.img, args.img, 80, args.dist, 1234 + rank, dev)
    anchors = torch.tensor(synth.V5_ANCHORS_PX) if args.family in ("yolov5", "yolov7") else None
    pp = PostProcessor(args.family, hyp, anchors=anchors)
    flat = flatten_heads(args.family, heads)
    ent = pp._prepare(flat, args.batch, args.img, args.img, _lib.INPUT_RAW_HEADS)
    lib = _lib.load()
    params, N, out = ent["params"], ent["N"], ent["out"]
    ptrs = _lib.head_pointer_array(flat)
    max_det = params.max_det
    n_row_f = args.batch * max_det * 6

    class Slot:
        """Intermediates + outputs of one in-flight batch.  `flat_send` = [rows (b, max_det, 6) f32 | counts (b) i32]:
        the NMS kernel writes straight into it and, at N > 1, it is the fixed-stride send buffer of the all-gather."""

        def __init__(self, send_row=None):
            self.keys = torch.empty((args.batch, N), dtype=torch.int64, device=dev)
            self.counts = torch.zeros((args.batch, 4), dtype=torch.int32, device=dev)
            self.flat_send = send_row if send_row is not None else torch.zeros(n_row_f + args.batch, dtype=torch.float32, device=dev)
            self.dets = self.flat_send[:n_row_f].view(args.batch, max_det, 6)
            self.cnt = self.flat_send[n_row_f:].view(torch.int32)
            self.idx = torch.empty((args.batch, max_det), dtype=torch.int32, device=dev)
            self.gathered = torch.empty((world, n_row_f + args.batch), dtype=torch.float32, device=dev) if world > 1 else None
            self.filtered = torch.cuda.Event()
            self.done = torch.cuda.Event()
            self.nms_done = torch.cuda.Event()

    # Two slots + two streams: the select/sort/NMS kernel of batch i (64 CTAs, latency-bound) runs on the side stream
    # while the HBM-bound filter kernel of batch i+1 streams on the main one.  --pipeline 0 serialises them.
    n_lanes = max(2, args.lanes) if args.pipeline == 3 else 2
    # mode 3: slot i % L lives on lane i % L.  At N > 1 the L send buffers are rows of ONE tensor, all-gathered once per
    # L steps (one NCCL launch per cycle instead of per step: the exchange is latency-bound, 461 kB per rank and step).
    big_send = torch.zeros((n_lanes, n_row_f + args.batch), dtype=torch.float32, device=dev) if args.pipeline == 3 else None
    big_recv = torch.empty((world, n_lanes, n_row_f + args.batch), dtype=torch.float32, device=dev) if (world > 1 and args.pipeline == 3) else None
    slots = ([Slot(big_send[i]) for i in range(n_lanes)] if args.pipeline == 3 else [Slot(), Slot()]) if args.pipeline else [Slot()]
    comm = torch.cuda.Stream(device=dev) if (world > 1 and args.pipeline == 3) else None
    gathered_ev = torch.cuda.Event()

    def gather_cycle():
        """all-gather of the L most recent batches' rows+counts on the comm stream"""
        for sl_ in slots:
            comm.wait_event(sl_.nms_done)
        with torch.cuda.stream(comm):
            dist.all_gather_into_tensor(big_recv.view(-1), big_send.view(-1))
        gathered_ev.record(comm)

    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev, priority=-1 if args.pipeline == 2 else 0) if args.pipeline else stream

    def launch_filter(head_ptrs, sl, st):
        _lib.check(lib.ysb_filter_candidates(ctypes.byref(params), head_ptrs, len(flat), sl.keys.data_ptr(), N,
                                             sl.counts.data_ptr(), ctypes.c_void_p(st.cuda_stream)),
                   "ysb_filter_candidates")

    def launch_nms(head_ptrs, sl, st):
        _lib.check(lib.ysb_select_nms(ctypes.byref(params), head_ptrs, len(flat), sl.keys.data_ptr(), N,
                                      sl.counts.data_ptr(), sl.dets.data_ptr(), sl.idx.data_ptr(),
                                      sl.cnt.data_ptr(), ctypes.c_void_p(st.cuda_stream)), "ysb_select_nms")

    step_no = [0]
    lanes = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)] if args.pipeline == 3 else None
    graphs = None
    if args.graph and args.pipeline == 3:
        # one graph per lane: {zero the counters, filter kernel, NMS kernel} with that lane's buffers baked in
        launch_filter(ptrs, slots[0], stream)
        launch_nms(ptrs, slots[0], stream)   # first launches outside capture (function attributes, lazy module load)
        torch.cuda.synchronize()
        graphs = []
        for i in range(n_lanes):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=lanes[i], capture_error_mode="thread_local"):
                launch_filter(ptrs, slots[i], lanes[i])
                launch_nms(ptrs, slots[i], lanes[i])
            graphs.append(g)

    def step(ev=None, head_ptrs=None):
        head_ptrs = head_ptrs or ptrs
        sl = slots[step_no[0] % len(slots)]
        if args.pipeline == 3:
            # two independent lanes: step i runs filter -> NMS (-> all-gather) in order on stream i % 2, so the NMS
            # kernel of one step overlaps the filter kernel of the next without any cross-stream event
            st = lanes[step_no[0] % n_lanes]
            if graphs is not None and head_ptrs is ptrs and comm is None:
                gi = step_no[0] % n_lanes
                step_no[0] += 1
                with torch.cuda.stream(st):
                    if ev:
                        ev[0].record(st)
                    graphs[gi].replay()
                    if ev:
                        for e_ in ev[1:]:
                            e_.record(st)
                return sl
            step_no[0] += 1
            if ev:
                ev[0].record(st)
            launch_filter(head_ptrs, sl, st)
            if ev:
                ev[1].record(st)
                ev[2].record(st)
            if comm is not None:
                st.wait_event(gathered_ev)   # the slot's send row may be overwritten only after its cycle was gathered
            launch_nms(head_ptrs, sl, st)
            if ev:
                ev[3].record(st)
            if comm is not None:
                sl.nms_done.record(st)
                if step_no[0] % n_lanes == 0:
                    gather_cycle()
            return sl
        step_no[0] += 1
        if args.pipeline:
            stream.wait_event(sl.done)       # the slot's previous batch has left the NMS stage
        if ev:
            ev[0].record(stream)
        launch_filter(head_ptrs, sl, stream)
        if ev:
            ev[1].record(stream)
        if args.pipeline:
            sl.filtered.record(stream)
            side.wait_event(sl.filtered)
        if ev:
            ev[2].record(side)
        launch_nms(head_ptrs, sl, side)
        if ev:
            ev[3].record(side)
        if world > 1:
            with torch.cuda.stream(side):
                dist.all_gather_into_tensor(sl.gathered.view(-1), sl.flat_send)
        if args.pipeline:
            sl.done.record(side)
        return sl

    def fork():
        if args.pipeline == 3:  # the lanes start after whatever the main stream has queued (e_beg, H2D copies)
            for ln in lanes:
                ln.wait_stream(stream)

    def drain():
        if args.pipeline == 3:
            if comm is not None:
                if step_no[0] % n_lanes != 0:   # a partial last cycle still has to be exchanged
                    gather_cycle()
                stream.wait_stream(comm)
            for ln in lanes:
                stream.wait_stream(ln)
        elif args.pipeline:
            stream.wait_stream(side)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    # ---- timed region: exactly K steps, device events, max over ranks -------------------------------------
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.arm(True)  # sampled through the timed region and the identical-load hold phase that follows it
    e_beg.record(stream)
    fork()
    for k in range(args.steps):
        step(evs[k])
    drain()
    e_end.record(stream)
    sync_all()
    total_ms = e_beg.elapsed_time(e_end)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    # Keep the identical load running ~1.5 s so NVML (10 ms period) sees the clocks this workload runs at.  The number
    # of extra steps is derived from the (all-reduced) step time, so every rank issues the same collectives.
    hold_steps = int(min(200000, max(100, 1500.0 / max(total_ms / args.steps, 1e-3))))
    if sampler:
        sampler.arm(True)
    for i in range(hold_steps):
        step()
        if i % 64 == 63:
            drain()
            torch.cuda.synchronize()
    drain()
    torch.cuda.synchronize()
    if sampler:
        sampler.arm(False)
    sync_all()
    filt_ms = [e[0].elapsed_time(e[1]) for e in evs]
    nms_ms = [e[2].elapsed_time(e[3]) for e in evs]
    # the dominant kernel alone (no concurrent NMS kernel), same inputs, same stream, CUDA events around each launch
    iso = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(50)]
    for a, b_ in iso:
        a.record(stream)
        launch_filter(ptrs, slots[0], stream)
        b_.record(stream)
    torch.cuda.synchronize()
    iso_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in iso)
    m_mean = float(slots[0].counts[:, 0].float().mean().item())
    # every rank's own speed on the HBM-bound kernel: the collectives make all ranks run at the slowest one's pace
    iso_per_rank = [iso_ms]
    if world > 1:
        tt = torch.tensor([iso_ms], dtype=torch.float64, device=dev)
        allv = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allv, tt)
        iso_per_rank = [float(x) for x in allv.tolist()]

    # ---- end-to-end: host (pinned) heads -> H2D -> kernels -> D2H of rows + counts, per step -------------------
    host_heads = [torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True).copy_(t_) for t_ in flat]
    dev_heads = [torch.empty_like(t_) for t_ in flat]
    host_dets = torch.empty((args.batch, max_det, 6), dtype=torch.float32, pin_memory=True)
    host_cnt = torch.empty((args.batch,), dtype=torch.int32, pin_memory=True)
    ptrs2 = _lib.head_pointer_array(dev_heads)

    def e2e_step():
        for d, h in zip(dev_heads, host_heads):
            d.copy_(h, non_blocking=True)
        fork()
        sl = step(None, ptrs2)
        drain()
        host_dets.copy_(sl.dets, non_blocking=True)
        host_cnt.copy_(sl.cnt, non_blocking=True)
        stream.synchronize()  # the caller reads the rows on the host after every call

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = sum(t_.numel() * 4 for t_ in flat)
    d2h = host_dets.numel() * 4 + host_cnt.numel() * 4

    clocks = sampler.stop() if sampler else None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        C = 80
        # channels the filter kernel must read per candidate: class logits (+ objectness / centerness / conf when the
        # family has one); box channels are not read by it
        n_read_ch = C + (0 if args.family in ("yolov8", "retinanet") else 1)
        algo_bytes = args.batch * (N * n_read_ch * 4 + m_mean * 8)
        filt_mean_ms = statistics.mean(filt_ms)
        # With several batches in flight the launches of the dominant kernel overlap each other and the NMS kernels, so
        # an event-delimited "launch duration" double-counts time.  Its average duration over the timed region is the
        # region itself divided by the launches it contains: K launches moved K * algo_bytes in total_ms.
        region_launch_ms = total_ms / args.steps if args.pipeline == 3 else filt_mean_ms
        achieved = algo_bytes / (region_launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * args.batch * args.steps / (total_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (+f64 IoU test)",
            "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": world * args.batch * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "host pinned head tensors -> ysb_filter_candidates + ysb_select_nms -> host rows/counts"},
            "gpu_launches": 2 * args.steps,
            "filter_alone_ms_per_rank": [round(x, 5) for x in iso_per_rank],
            "roofline": {"kernel": ("k_filter_rows" if args.family in ("yolov7", "retinanet", "retinanet_exp") else "k_filter_planes_v4<9,128,6>" if args.family in ("yolov5", "yolox", "yolov8") else "k_filter_planes<4>")
                                   + " (decode-sigmoid + filter + class pick + compaction)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         "traffic": NCU_TRAFFIC_BYTES.get((args.family, args.img, args.batch, args.dist)),
                         "traffic_source": "profiles/r1_filter_ncu_raw.txt (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture)",
                         "launch_ms_alone": iso_ms, "achieved_alone": algo_bytes / (iso_ms * 1e-3) / 1e9,
                         "frac_alone": algo_bytes / (iso_ms * 1e-3) / 1e9 / peak,
                         "note": "achieved/frac: algorithmic bytes of the K filter launches / duration of the timed region "
                                 "(launches of consecutive batches overlap each other and the NMS kernels, so this is a "
                                 "lower bound for the kernel); *_alone: the same kernel timed by itself with CUDA events; "
                                 "launch_ms_overlapped: event-delimited duration of one launch inside the region",
                         "launch_ms_overlapped": filt_mean_ms,
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "bytes_per_image": f"N*{n_read_ch}*4 read + M*8 written = {N * n_read_ch * 4} + {m_mean * 8:.0f}",
                         "launch_ms": region_launch_ms},
            "pipeline": bool(args.pipeline), "lanes": n_lanes if args.pipeline == 3 else (2 if args.pipeline else 1),
            "cuda_graph": bool(graphs),
            "stages_ms": {"filter_compact": filt_mean_ms, "select_sort_nms": statistics.mean(nms_ms),
                          "filter_p50": statistics.median(filt_ms), "nms_p50": statistics.median(nms_ms)},
            "survivors_per_image": m_mean,
        }
        if not args.no_cpu_baseline and world == 1 and args.family == "yolov5":
            v, dt = cpu_images_per_second(cpu_heads(args.cpu_images, args), 1)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"{args.cpu_images} images of the same workload on 1 host thread ({dt:.1f} s): C/numpy port "
                          "of the reference algorithm incl. the full greedy NMS loop (no early stop)"}
            line["parity"] = parity_counters(args, pp, heads, slots[0], hyp)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
