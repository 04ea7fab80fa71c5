"""The N>1 path of the PRODUCT on the GPU box the driver has: two ranks share cuda:0 (NCCL refuses two ranks on
one device, so the collective runs over gloo on CUDA tensors); each rank plays its contiguous shard of global
game ids with the CUDA kernels, the statistics are all-reduced (simulator.batch.all_reduce_stats), and the result
must be the single-rank answer and the oracle's.  The NCCL variant of the same check runs
inside bench.py --gpus N (`stats_equal_to_single_rank`)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import DEFAULT_BOUNCE_GRID, PRODUCT, ROOT

pytestmark = pytest.mark.gpu
N_TOTAL = 200_003


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
    from simulator import batch

    torch.cuda.set_device(0)
    start, count = batch.shard_range(N_TOTAL, rank, world)
    res = batch.connect_rollout((6, 7, 4), count, 3, start, per_game=True)
    batch.all_reduce_stats(res.stats)
    bres = batch.bounce_rollout(np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8), count, 3, start, max_plies=256)
    batch.all_reduce_stats(bres.stats)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"stats{rank}.npy"), res.stats.cpu().numpy())
    np.save(os.path.join(out_dir, f"len{rank}.npy"), res.length.cpu().numpy())
    np.save(os.path.join(out_dir, f"bstats{rank}.npy"), bres.stats.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_of_the_product_reduce_to_the_single_rank_answer(tmp_path, oracle):
    from simulator import batch

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = batch.connect_rollout((6, 7, 4), N_TOTAL, 3, 0, per_game=True)
    bsingle = batch.bounce_rollout(np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8), N_TOTAL, 3, 0, max_plies=256)
    torch.cuda.synchronize()
    s0, s1 = np.load(tmp_path / "stats0.npy"), np.load(tmp_path / "stats1.npy")
    np.testing.assert_array_equal(s0, s1)
    np.testing.assert_array_equal(s0, single.stats.cpu().numpy())
    lens = np.concatenate([np.load(tmp_path / "len0.npy"), np.load(tmp_path / "len1.npy")])
    np.testing.assert_array_equal(lens, single.length.cpu().numpy())
    np.testing.assert_array_equal(np.load(tmp_path / "bstats0.npy"), bsingle.stats.cpu().numpy())
    ref = oracle.connect_rollout(6, 7, 4, N_TOTAL, gid0=0, seed=3, want_actions=False, want_grid=False)
    np.testing.assert_array_equal(s0, ref["stats"])
