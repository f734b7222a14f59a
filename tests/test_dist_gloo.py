"""Multi-rank host logic on CPU: image sharding + the fixed-stride all-gather of kept detections (gloo, world_size 2/3)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yoloseries_b200.dist import gather_detections, shard_bounds


def test_shard_bounds_cover_batch():
    for batch in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_shard(batch, max_det, rank, world):
    g = torch.Generator().manual_seed(1234)
    dets = torch.rand((batch, max_det, 6), generator=g)
    cnt = torch.randint(-1, max_det + 1, (batch,), generator=g, dtype=torch.int32)
    lo, hi = shard_bounds(batch, rank, world)
    return dets, cnt, dets[lo:hi].clone(), cnt[lo:hi].clone()


def _worker(rank, world, port, batch, max_det):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full_d, full_c, my_d, my_c = _fake_shard(batch, max_det, rank, world)
        got_d, got_c = gather_detections(my_d, my_c, batch)
        assert got_d.shape == full_d.shape and got_c.shape == full_c.shape
        assert torch.equal(got_d, full_d), "rows differ after the all-gather"
        assert torch.equal(got_c, full_c), "counts (incl. the -1 'None' marker) differ after the all-gather"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 8), (2, 7), (3, 8)])
def test_gather_detections_gloo(world, batch):
    port = 29500 + (os.getpid() % 500) + world * 7 + batch
    mp.spawn(_worker, args=(world, port, batch, 12), nprocs=world, join=True)


def test_default_lanes_follow_batch_and_world():
    """Host-side choice of the batches in flight (dist.default_lanes): 4 on one GPU, 6 with peers, 8 below 64 images."""
    from yoloseries_b200.dist import default_lanes
    assert default_lanes(64, 1) == 4 and default_lanes(256, 1) == 4
    assert default_lanes(64, 2) == 6 and default_lanes(64, 8) == 6
    assert [default_lanes(b, w) for b in (1, 8, 16, 32, 63) for w in (1, 8)] == [8] * 10
