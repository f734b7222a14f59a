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
