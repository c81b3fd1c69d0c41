"""CPU, world_size 2 over gloo: the N>1 host logic (contiguous batch shards, no data-path collective,
max-over-ranks timing, whole-job throughput) used by bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rethink_acoustic_image_enhancement_b200.sharding import job_throughput, max_over_ranks, shard_slice


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    batch = torch.arange(13 * 4, dtype=torch.float32).view(13, 4)          # 13 "images"
    mine = batch[shard_slice(13, rank, world)]
    # each rank processes only its own units; the only cross-rank traffic is this check and the timing reductions
    s = mine.sum().clone()
    dist.all_reduce(s)
    t = max_over_ranks(10.0 + 5.0 * rank, dev)
    thr, ms = job_throughput(units_per_rank=8, steps=3, elapsed_ms_this_rank=100.0 * (rank + 1), device=dev)
    out[rank] = (mine.shape[0], float(s), t, thr, ms)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] == 7 and res[1][0] == 6                                  # 13 units -> 7 + 6, contiguous
    total = float(torch.arange(13 * 4, dtype=torch.float32).sum())
    assert res[0][1] == total and res[1][1] == total                          # every unit processed exactly once
    assert res[0][2] == 15.0 and res[1][2] == 15.0                            # max over ranks
    assert abs(res[0][3] - (2 * 8 * 3) / 0.2) < 1e-9 and res[0][4] == 200.0   # whole-job units / slowest rank


def test_shard_slices_partition_the_batch():
    for n in (1, 7, 64, 128, 1024):
        for world in (1, 2, 4, 8):
            covered = []
            for r in range(world):
                sl = shard_slice(n, r, world)
                covered += list(range(n))[sl]
            assert covered == list(range(n))
    assert shard_slice(64, 3, 8) == slice(24, 32)
