"""Multi-process host logic on CPU (gloo, world_size 2): contiguous image sharding + the final fixed-width gather
reproduce the single-process table in image order (the N>1 path of bench.py)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vpho_b200.distributed import RECORD_WIDTH, gather_records, image_record, shard_range


def _table(n):
    g = torch.Generator().manual_seed(0)
    return torch.randn(n, 21, 3, generator=g), torch.randn(n, 9, generator=g, dtype=torch.float64)


def _worker(rank, world, n_images, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        j, o = _table(n_images)
        lo, hi = shard_range(n_images, world, rank)
        full = gather_records(image_record(j[lo:hi], o[lo:hi]), n_images)
        ret[rank] = full
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_ranges_partition_the_images():
    for n, w in ((512, 8), (7, 2), (5, 4), (3, 8)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_gather_matches_single_process():
    n_images, world = 7, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, n_images, 29533, ret), nprocs=world, join=True)
    j, o = _table(n_images)
    ref = image_record(j, o)
    assert ref.shape == (n_images, RECORD_WIDTH)
    for r in range(world):
        assert torch.equal(ret[r], ref)
