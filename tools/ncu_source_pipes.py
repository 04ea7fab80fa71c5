"""Where do a kernel's instructions go -- by pipe, by source line, and at how many active lanes?

    ncu --set full --import-source on -k regex:<kernel> -c 1 -o rep <command>          (on the GPU box)
    ncu -i rep.ncu-rep --page source --csv --print-source sass > source.csv            (anywhere)
    cuobjdump -xelf all <object>.o && nvdisasm -g -c <object>.sm_100a.cubin > all.txt   (same build, -lineinfo)
    # cut the kernel's `.text.<mangled name>:` section out of all.txt into listing.txt, then
    python tools/ncu_source_pipes.py source.csv listing.txt [top]

The source page has the executed count of every SASS instruction; the nvdisasm listing maps instruction offsets to
source lines.  Every instruction is put into the pipe it issues to (ALU: LOP3 / SHF / ISETP / IADD3 / SEL / PRMT / LEA;
FMA: IMAD and its MOV / SHL / IADD forms; XU: POPC / FLO / BREV; LSU; control), and the table lists, per source line,
its share of all warp-instructions, the split by pipe, the average number of active lanes and its share of the
shared-memory wavefronts.  This is how the round's last steps were found: the Bounce segment held 42 % of the
ALU-pipe instructions of a kernel whose ALU pipe was saturated, and the line kernels spent 13 % of
their instructions and half of their shared-memory wavefronts clearing line words at ~4 active lanes.
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
listing = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40

cur, line_of = None, {}
for ln in open(listing):
    m = re.search(r'//## File "[^"]*/([A-Za-z_0-9]+\.(?:cuh?|hpp|h))", line (\d+)(.*)', ln)
    if m:
        if "inlined" not in m.group(3):
            cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur

FMA = {"IMAD", "FFMA", "FMUL", "FADD"}
XU = {"POPC", "FLO", "BREV", "MUFU", "I2F", "F2I"}
LSU = {"LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "REDG", "LDC", "LDL", "STL", "SHFL", "ATOM", "LDCU", "REDUX", "MATCH"}
CTL = {"BRA", "BSSY", "BSYNC", "BREAK", "EXIT", "WARPSYNC", "CALL", "RET", "NOP", "BAR", "VOTE", "VOTEU", "S2R", "S2UR",
       "R2UR", "UMOV", "UIADD3", "ULOP3", "USHF", "ULEA", "UISETP", "USEL", "UIMAD", "UPOPC", "UFLO", "R2P", "P2R", "DEPBAR",
       "ERRBAR", "MEMBAR", "CCTL", "BMOV", "NANOSLEEP", "UPLOP3", "UPRMT", "UBREV", "ELECT"}


def pipe(op):
    b = op.split(".")[0]
    return "FMA" if b in FMA else "XU" if b in XU else "LSU" if b in LSU else "CTL" if b in CTL else "ALU"


hdr = rows[1]
ex_i, th_i = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
wf_i = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
base, tot, ops = None, collections.Counter(), collections.Counter()
by_line = collections.defaultdict(collections.Counter)
lanes = collections.defaultdict(lambda: [0, 0])
wf = collections.Counter()
for r in rows[2:]:
    try:
        addr = int(r[0], 16)
    except (ValueError, IndexError):
        continue
    base = addr if base is None else base
    op = re.sub(r"^@!?U?P\d+\s+", "", r[1].strip()).split()[0]
    ex, th, pp = int(r[ex_i]), int(r[th_i]), pipe(op)
    key = line_of.get(addr - base)
    tot[pp] += ex
    ops[op.split(".")[0]] += ex
    by_line[key][pp] += ex
    lanes[key][0] += ex
    lanes[key][1] += th
    if wf_i is not None and r[wf_i]:
        try:
            wf[key] += int(r[wf_i])
        except ValueError:
            pass
T, W = sum(tot.values()), sum(wf.values()) or 1
print("warp-instructions", T, {k: round(100 * v / T, 1) for k, v in tot.items()}, "shared-memory wavefronts", W)
print("opcodes", [(k, round(100 * v / T, 1)) for k, v in ops.most_common(16)])
for key, c in sorted(by_line.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    s = sum(c.values())
    print(f"{str(key):34s} {100 * s / T:5.2f}%  ALU {100 * c['ALU'] / T:5.2f}  FMA {100 * c['FMA'] / T:5.2f}  LSU {100 * c['LSU'] / T:5.2f}  "
          f"CTL {100 * c['CTL'] / T:5.2f}  lanes {lanes[key][1] / max(lanes[key][0], 1):5.1f}  wavefronts {100 * wf[key] / W:5.1f}%")
