"""N>1 host-side logic on CPU: contiguous game-id shards + one all-reduce(sum) of the statistics
vector (gloo, world_size 2).  The per-rank "engine" here is the oracle (no GPU in the CPU suite); the same
check with the PRODUCT as the engine is tests/test_gpu_distributed.py (two ranks on one GPU, gloo on CUDA tensors)
and bench.py --gpus N (`stats_equal_to_single_rank`, NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PRODUCT, ROOT

N_TOTAL = 3001


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PRODUCT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as o
    from simulator.batch import all_reduce_stats, shard_range

    start, count = shard_range(N_TOTAL, rank, world)
    res = o.connect_rollout(6, 7, 4, count, gid0=start, seed=3, want_actions=False, want_grid=False)
    stats = torch.from_numpy(res["stats"].copy())
    all_reduce_stats(stats)
    np.save(os.path.join(out_dir, f"stats{rank}.npy"), stats.numpy())
    np.save(os.path.join(out_dir, f"len{rank}.npy"), res["length"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shards_reduce_to_the_single_rank_answer(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full = oracle.connect_rollout(6, 7, 4, N_TOTAL, gid0=0, seed=3, want_actions=False, want_grid=False)
    s0 = np.load(tmp_path / "stats0.npy")
    s1 = np.load(tmp_path / "stats1.npy")
    np.testing.assert_array_equal(s0, s1)
    np.testing.assert_array_equal(s0, full["stats"])
    lens = np.concatenate([np.load(tmp_path / "len0.npy"), np.load(tmp_path / "len1.npy")])
    np.testing.assert_array_equal(lens, full["length"])
