"""BASELINE.json configs[3]: rollout + trajectory + final grid + reward export on the larger boards.

    python tools/time_export.py [--games N] [--launches L] [--only H W K]

Prints, per board, the device time of the plain rollout (length + winner only), of the full export and
of the partial exports, and the export's share as GB/s of algorithmic output bytes against the HBM copy peak.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))

import torch  # noqa: E402

from simulator import _native as N  # noqa: E402
from simulator import batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4 * 2**20)
ap.add_argument("--launches", type=int, default=7)
ap.add_argument("--only", nargs=3, type=int, default=None)
ap.add_argument("--lib", default=None, help="time another build of libbgs_b200.so (kernel experiments)")
args = ap.parse_args()
if args.lib:
    N.LIB_PATH = os.path.abspath(args.lib)
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
hbm = json.load(open(peaks)).get("hbm_gbs", 6444.4) if os.path.exists(peaks) else 6444.4
stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device="cuda")
n = args.games


def timed(cfg, **kw):
    res, ms, steps = None, [], []
    for i in range(args.launches + 2):
        stats.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        res = batch.connect_rollout(cfg, n, 1, i * n, per_game=True, stats=stats, out=res, **kw)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(a.elapsed_time(b))
            steps.append(int(stats[N.STAT_STEPS]))
    return statistics.median(ms), statistics.mean(steps)


boards = [tuple(args.only)] if args.only else [(8, 9, 5), (10, 12, 6)]
for cfg in boards:
    hw = cfg[0] * cfg[1]
    ms0, st0 = timed(cfg)
    out = {"board": cfg, "games": n, "plain_ms": round(ms0, 4), "plain_Gsteps": round(st0 / ms0 / 1e6, 1)}
    for name, kw, nbytes in (("full", dict(actions=True, final_grid=True, reward=True), 2 * hw + 10),
                             ("actions", dict(actions=True), hw + 2),
                             ("grid", dict(final_grid=True), hw + 2),
                             ("reward", dict(reward=True), 10)):
        ms1, _ = timed(cfg, **kw)
        out[name + "_ms"] = round(ms1, 4)
        if name == "full":
            out["export_ms"] = round(ms1 - ms0, 4)
            out["export_GBps"] = round(nbytes * n / (ms1 - ms0) / 1e6, 1)
            out["export_frac_of_hbm_copy_peak"] = round(nbytes * n / (ms1 - ms0) / 1e6 / hbm, 3)
            out["whole_launch_GBps"] = round(nbytes * n / ms1 / 1e6, 1)
    print(json.dumps(out), flush=True)
