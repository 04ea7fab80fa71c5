"""Small batches of every kernel, for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from simulator import batch  # noqa: E402

for cfg, n in (((6, 7, 4), 5000), ((8, 9, 5), 1500), ((10, 12, 6), 1000), ((2, 3, 2), 700), ((5, 5, 3), 700), ((7, 8, 4), 700)):
    r = batch.connect_rollout(cfg, n, 3, 11, per_game=True, actions=True, final_grid=True, reward=True)
    torch.cuda.synchronize()
    print(cfg, r.stats_dict())
b = batch.ConnectBatch.initial((6, 7, 4), 777)
for i in range(10):
    b, st = b.step(torch.randint(-1, 8, (777,)))
torch.cuda.synchronize()
grid = np.zeros((9, 6), dtype=np.int8)
grid[1] = grid[7] = [1, 2, 3, 3, 2, 1]
r = batch.bounce_rollout(grid, 3000, 5, 0, max_plies=64, moves=True, final_grid=True, reward=True)
torch.cuda.synchronize()
print("bounce", r.stats_dict())
bb = batch.BounceBatch.initial(grid, 333)
row, targets, count = bb.moves()
mv = torch.zeros((333, 4), dtype=torch.int32)
mv[:, 0] = 1; mv[:, 1] = 1; mv[:, 2] = 1; mv[:, 3] = 3
bb, st = bb.step(mv)
torch.cuda.synchronize()
print("ok", int(st.sum()))
