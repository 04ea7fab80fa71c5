"""Small driver for ncu: a few launches of the Connect rollout kernel (and the export path).

    python tools/profile_connect.py [H W K] [--games N] [--launches L] [--export]
"""
import argparse
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))

import torch  # noqa: E402

from simulator import batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("cfg", nargs="*", type=int, default=[6, 7, 4])
ap.add_argument("--games", type=int, default=16 * 2**20)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--export", action="store_true")
args = ap.parse_args()
cfg = tuple(args.cfg)
res = None
for i in range(args.launches):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    res = batch.connect_rollout(cfg, args.games, 1, i * args.games, per_game=True, actions=args.export,
                                final_grid=args.export, reward=args.export, out=res)
    b.record()
    torch.cuda.synchronize()
    s = res.stats_dict()
    print(f"launch {i}: {a.elapsed_time(b):.3f} ms, {s['steps']} env-steps, {s['steps'] / a.elapsed_time(b) / 1e6:.2f} G steps/s")
