"""One-off evidence run: every game of a full-size batch (16 Mi by default) is replayed through the
oracle's State/Action transition on the host (all cores) and must agree bit for bit -- moves legal,
game ended exactly at `length`, winner, final grid and reward identical.

    python tools/replay_check.py [H W K] [--games N] [--seed S]
"""
import argparse
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import binding as oracle  # noqa: E402  (the checker)
from simulator import batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("cfg", nargs="*", type=int, default=[6, 7, 4])
ap.add_argument("--games", type=int, default=16 * 2**20)
ap.add_argument("--seed", type=int, default=20261018)
args = ap.parse_args()
H, W, K = args.cfg
t0 = time.perf_counter()
res = batch.connect_rollout((H, W, K), args.games, args.seed, 0, per_game=True, actions=True, final_grid=True, reward=True)
torch.cuda.synchronize()
t_gpu = time.perf_counter() - t0
acts, length, winner = res.actions.cpu().numpy(), res.length.cpu().numpy(), res.winner.cpu().numpy()
grid, reward = res.final_grid.cpu().numpy(), res.reward.cpu().numpy()
oracle.lib()
workers = os.cpu_count() or 1
chunks = np.array_split(np.arange(args.games), workers * 4)


def check(ix):
    lo, hi = int(ix[0]), int(ix[-1]) + 1
    return oracle.connect_replay(H, W, K, acts[lo:hi], length[lo:hi], winner[lo:hi], grid[lo:hi], reward[lo:hi])[0]


t0 = time.perf_counter()
with ThreadPoolExecutor(workers) as ex:
    bad = sum(ex.map(check, chunks))
t_cpu = time.perf_counter() - t0
s = res.stats_dict()
print(f"Connect({H},{W},{K}) seed={args.seed}: {args.games} games / {s['steps']} env-steps generated on the GPU in {t_gpu * 1e3:.1f} ms "
      f"(incl. export + launch), replayed through the oracle on {workers} host threads in {t_cpu:.1f} s: "
      f"{bad} mismatching games ({100.0 * (args.games - bad) / args.games:.4f} % agreement); "
      f"p0/p1/draw = {s['wins0']}/{s['wins1']}/{s['draws']}, mean length {s['steps'] / s['games']:.3f}")
sys.exit(1 if bad else 0)
