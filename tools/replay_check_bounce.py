"""One-off evidence run for Bounce: 1 Mi default games replayed through the oracle on all host cores."""
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

n, cap = 2**20, 512
grid = np.zeros((9, 6), dtype=np.int8)
grid[1] = grid[7] = [1, 2, 3, 3, 2, 1]  # reference src/simulator/textual/bounce.py:66-78
res = batch.bounce_rollout(grid, n, 20261018, 0, max_plies=cap, moves=True, final_grid=True, reward=True)
torch.cuda.synchronize()
moves, length, winner = res.actions.cpu().numpy(), res.length.cpu().numpy().astype(np.uint16), res.winner.cpu().numpy()
fgrid, reward = res.final_grid.cpu().numpy(), res.reward.cpu().numpy()
oracle.lib()
workers = os.cpu_count() or 1
chunks = np.array_split(np.arange(n), workers * 8)


def check(ix):
    lo, hi = int(ix[0]), int(ix[-1]) + 1
    return oracle.bounce_replay(grid, moves[lo:hi], length[lo:hi], winner[lo:hi], fgrid[lo:hi], reward[lo:hi])[0]


t0 = time.perf_counter()
with ThreadPoolExecutor(workers) as ex:
    bad = sum(ex.map(check, chunks))
s = res.stats_dict()
print(f"Bounce default 9x6 seed=20261018 max_plies={cap}: {n} games / {s['steps']} env-steps from the GPU replayed through "
      f"the oracle on {workers} host threads in {time.perf_counter() - t0:.1f} s: {bad} mismatching games "
      f"({100.0 * (n - bad) / n:.4f} % agreement); p0/p1/draw/truncated = {s['wins0']}/{s['wins1']}/{s['draws']}/{s['truncated']}, "
      f"mean length {s['steps'] / s['games']:.3f}, max length {int(length.max())}")
sys.exit(1 if bad else 0)
