"""Timing of the batched single-step kernels (row f1): ConnectBatch.step / query, BounceBatch.moves / step."""
import os
import statistics
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from simulator import batch  # noqa: E402
from simulator import _native as _N  # noqa: E402

if len(sys.argv) > 2 and sys.argv[1] == "--lib":  # time another build of libbgs_b200.so (kernel experiments)
    _N.LIB_PATH = os.path.abspath(sys.argv[2])


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return statistics.median(ms)


for cfg in ((6, 7, 4), (10, 12, 6)):
    n = 4 * 2**20
    H, W, K = cfg
    b = batch.ConnectBatch.initial(cfg, n)
    # a mid-game batch: play 10 random plies
    for _ in range(10):
        acts = torch.randint(0, W, (n,), device="cuda")
        b, _ = b.step(acts)
    acts = torch.randint(0, W, (n,), device="cuda", dtype=torch.int32)
    ms = timed(lambda: b.step(acts))
    bytes_moved = n * (2 * H * W + 1 + 1 + 4 + 1 + 1 + 1 + 8 + 4 + 4)
    print(f"connect {cfg} step n={n}: {ms:.3f} ms, {n / ms / 1e6:.2f} G states/s, {bytes_moved / ms / 1e6:.0f} GB/s algorithmic")
    ms = timed(lambda: b._query())
    print(f"connect {cfg} query n={n}: {ms:.3f} ms, {n * (H * W + 1 + 1 + 4 + 8) / ms / 1e6:.0f} GB/s algorithmic")

grid = np.zeros((9, 6), dtype=np.int8)
grid[1] = grid[7] = [1, 2, 3, 3, 2, 1]
n = 2**20
bb = batch.BounceBatch.initial(grid, n)
ms = timed(lambda: bb.moves(), 5)
print(f"bounce moves n={n}: {ms:.3f} ms, {n / ms / 1e6:.3f} G states/s")
row, targets, count = bb.moves()
mv = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
mv[:, 0] = 1; mv[:, 1] = 1; mv[:, 2] = 1; mv[:, 3] = 3
ms = timed(lambda: bb.step(mv), 5)
print(f"bounce step n={n}: {ms:.3f} ms, {n / ms / 1e6:.3f} G states/s")
