"""Timing of the per-ply grid export (HBM-write bound: (H*W+1)*H*W bytes per game)."""
import os
import statistics
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))

import torch  # noqa: E402

from simulator import _native as N  # noqa: E402
from simulator import batch  # noqa: E402

if len(sys.argv) > 2 and sys.argv[1] == "--lib":  # time another build of libbgs_b200.so (kernel experiments)
    N.LIB_PATH = os.path.abspath(sys.argv[2])

for cfg, n in (((6, 7, 4), 2**20), ((8, 9, 5), 2**18), ((10, 12, 6), 2**17)):
    H, W, K = cfg
    res = batch.connect_rollout(cfg, n, 1, 0, per_game=True, actions=True)
    out = batch.connect_trajectory_grids(cfg, res.actions, res.length)
    ms = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        batch.connect_trajectory_grids(cfg, res.actions, res.length, out=out)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    m = statistics.median(ms)
    nbytes = n * (H * W + 1) * H * W + n * (H * W + 1)
    print(f"trajectory grids {cfg} n={n}: {m:.3f} ms, {nbytes / m / 1e6:.0f} GB/s algorithmic "
          f"({nbytes / 1e6:.0f} MB), {int(res.length.sum()) / m / 1e6:.2f} G positions/s")
