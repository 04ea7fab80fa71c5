"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump: executed warp-instructions per
source line (and per SASS range), per kernel.   python tools/ncu_source_hot.py dump.csv [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
kern, hdr, per = None, None, {}
cur_file = None
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        cur_file = r[1]
    elif r and r[0] == "Function Name":
        kern = r[1]
    elif r and r[0] == "Line No":
        hdr = r
    elif r and hdr and len(r) == len(hdr) and kern:
        d = dict(zip(hdr, r))
        if not d["Line No"].strip():  # a SASS row: already counted in its source line's row
            i += 1
            continue
        try:
            ex = int(d["Instructions Executed"])
        except Exception:
            i += 1
            continue
        key = (kern, cur_file.split("/")[-1] if cur_file else "?", d["Line No"])
        per.setdefault(kern, defaultdict(lambda: [0, 0, ""]))
        e = per[kern][key[1:]]
        e[0] += ex
        try:
            e[1] += int(d["Thread Instructions Executed"])
        except Exception:
            pass
    i += 1
for k, m in per.items():
    tot = sum(v[0] for v in m.values())
    print("==", k[:110], "total warp-instr", tot)
    for (f, line), v in sorted(m.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {f}:{line:>5}  {v[0]:>12}  {100.0 * v[0] / tot:5.1f}%  lanes {v[1] / max(v[0], 1):5.1f}")
