"""Kernel-only timing of the Connect rollout: warm-up, then N launches timed with CUDA events.

    python tools/time_rollout.py [H W K] [--games N] [--launches L] [--actions] [--grid]
"""
import argparse
import os
import statistics
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))

import torch  # noqa: E402

from simulator import _native as N  # noqa: E402
from simulator import batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("cfg", nargs="*", type=int, default=[6, 7, 4])
ap.add_argument("--games", type=int, default=16 * 2**20)
ap.add_argument("--launches", type=int, default=10)
ap.add_argument("--actions", action="store_true")
ap.add_argument("--grid", action="store_true")
ap.add_argument("--lib", default=None, help="time another build of libbgs_b200.so (kernel experiments)")
args = ap.parse_args()
if args.lib:
    N.LIB_PATH = os.path.abspath(args.lib)
cfg = tuple(args.cfg)
stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device="cuda")
res = None
ms, steps = [], []
for i in range(args.launches + 3):
    stats.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    res = batch.connect_rollout(cfg, args.games, 1, i * args.games, per_game=True, actions=args.actions,
                                final_grid=args.grid, reward=args.grid, stats=stats, out=res)
    b.record()
    torch.cuda.synchronize()
    if i >= 3:
        ms.append(a.elapsed_time(b))
        steps.append(int(stats[N.STAT_STEPS]))
med = statistics.median(ms)
print(f"{cfg} games={args.games} actions={args.actions} grid={args.grid} lib={os.path.basename(os.path.dirname(N.LIB_PATH)) if args.lib else 'default'}: "
      f"ms min/med/max = {min(ms):.3f}/{med:.3f}/{max(ms):.3f}; {statistics.mean(steps) / med / 1e6:.1f} G env-steps/s (median)")
