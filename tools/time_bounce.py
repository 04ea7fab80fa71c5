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
ap.add_argument("--streams", type=int, default=0, help="also time --batches back-to-back batches alternating over this many streams")
ap.add_argument("--batches", type=int, default=8)
ap.add_argument("--lib", default=None, help="time another build of libbgs_b200.so (kernel experiments)")
args = ap.parse_args()
if args.lib:
    N.LIB_PATH = os.path.abspath(args.lib)
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

if args.streams > 0:
    # Sustained throughput over back-to-back batches: on one stream the straggler tail of every batch
    # (one ~400-ply game among 4 Mi is a 1.4 ms dependent chain) leaves the GPU nearly idle; with the
    # batches alternating over several streams the next batch's CTAs fill the SMs as the previous
    # batch's CTAs exit.
    streams = [torch.cuda.Stream() for _ in range(args.streams)]
    bstats = torch.zeros((args.batches, N.STATS_LEN), dtype=torch.int64, device="cuda")
    for rep in range(2):
        bstats.zero_()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for st in streams:
            st.wait_event(t0)
        done = []
        for b in range(args.batches):
            st = streams[b % args.streams]
            with torch.cuda.stream(st):
                batch.bounce_rollout(grid, args.games, 1, (100 + b) * args.games, max_plies=args.max_plies, stats=bstats[b])
                e = torch.cuda.Event()
                e.record()
                done.append(e)
        for e in done:
            torch.cuda.current_stream().wait_event(e)
        t1.record()
        torch.cuda.synchronize()
    total_ms = t0.elapsed_time(t1)
    total_steps = int(bstats[:, N.STAT_STEPS].sum())
    print(f"bounce {args.batches} batches of {args.games} games over {args.streams} stream(s): {total_ms:.2f} ms total, "
          f"{total_ms / args.batches:.2f} ms per batch, {total_steps / total_ms / 1e6:.2f} G env-steps/s sustained")
