"""Kernel-only timing of the Bounce rollout on run-time geometries (boards other than the default 9x6).

    python tools/time_bounce_boards.py [--lib other/libbgs_b200.so]
"""
import os
import statistics
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from simulator import _native as N  # noqa: E402
from simulator import batch  # noqa: E402

if len(sys.argv) > 2 and sys.argv[1] == "--lib":
    N.LIB_PATH = os.path.abspath(sys.argv[2])


def board(H, W, vals):
    g = np.zeros((H, W), dtype=np.int8)
    g[1] = g[H - 2] = [vals[x % len(vals)] for x in range(W)]
    return g


n = 2 * 2**20
stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device="cuda")
for name, g in (("6x3", board(6, 3, [1, 2, 3])), ("8x7", board(8, 7, [1, 2, 3, 3, 2, 1, 2])), ("7x5", board(7, 5, [1, 2, 3, 2, 1])),
                ("9x6 values<=7", board(9, 6, [1, 2, 7, 5, 2, 1]))):
    ms, steps = [], []
    for i in range(7):
        stats.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        batch.bounce_rollout(g, n, 1, i * n, max_plies=512, stats=stats)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(a.elapsed_time(b))
            steps.append(int(stats[N.STAT_STEPS]))
    med = statistics.median(ms)
    print(f"bounce {name} games={n}: {med:.2f} ms, {statistics.mean(steps) / med / 1e6:.2f} G env-steps/s")
