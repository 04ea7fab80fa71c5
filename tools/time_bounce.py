"""Kernel-only timing of the Bounce rollout on the default 9x6 board (BASELINE.json configs[2])."""
import argparse
import os
import statistics
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from simulator import _native as N  # noqa: E402
from simulator import batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4 * 2**20)
ap.add_argument("--launches", type=int, default=5)
ap.add_argument("--max-plies", type=int, default=512)
args = ap.parse_args()
grid = np.zeros((9, 6), dtype=np.int8)
grid[1] = grid[7] = [1, 2, 3, 3, 2, 1]  # reference src/simulator/textual/bounce.py:66-78
stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device="cuda")
ms, steps = [], []
for i in range(args.launches + 2):
    stats.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    res = batch.bounce_rollout(grid, args.games, 1, i * args.games, max_plies=args.max_plies, stats=stats)
    b.record()
    torch.cuda.synchronize()
    if i >= 2:
        ms.append(a.elapsed_time(b))
        steps.append(int(stats[N.STAT_STEPS]))
s = res.stats_dict()
med = statistics.median(ms)
print(f"bounce default 9x6 games={args.games} max_plies={args.max_plies}: ms min/med/max = {min(ms):.2f}/{med:.2f}/{max(ms):.2f}; "
      f"{statistics.mean(steps) / med / 1e6:.2f} G env-steps/s; mean plies {s['steps'] / s['games']:.2f}; "
      f"p0/p1/draw/truncated = {s['wins0']}/{s['wins1']}/{s['draws']}/{s['truncated']}")
