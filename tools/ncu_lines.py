"""Per-function and per-source-line executed warp-instructions per thread from an ncu source page in the combined view
(ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv).   python tools/ncu_lines.py X.csv <warps> [file]"""
import collections
import csv
import os
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
W = float(sys.argv[2]) if len(sys.argv) > 2 else 2048.0
SRC = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym_art_b200", "csrc", "qs_device.cuh")
hdr = None
per_line = collections.OrderedDict()
cur = None
for r in rows:
    if "Instructions Executed" in r:
        hdr = r
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= iE:
        continue
    if r[0].strip().isdigit():
        cur = (int(r[0]), r[1].strip()[:110])
        per_line.setdefault(cur, [0, 0])
    elif r[2].startswith("0x") and cur is not None:
        try:
            per_line[cur][0] += int(r[iE])
            per_line[cur][1] += int(r[iS])
        except ValueError:
            pass
tot = sum(v[0] for v in per_line.values())
print("total warp-instructions per thread: %.1f" % (tot / W))
src = open(SRC).read().split("\n")
fn_at, cur_fn = [None] * (len(src) + 2), "?"
for i, l in enumerate(src, 1):
    m = re.search(r"\b(\w+)\s*\((?:const |St<|Params|R\b|U4|uint|unsigned|long|int|double|float|void|bool|char)", l)
    if m and ("__device__" in l or "__global__" in l or (i > 1 and ("__device__" in src[i - 2] or "__global__" in src[i - 2]))) and "asm" not in l:
        cur_fn = m.group(1)
    fn_at[i] = cur_fn
agg, smp = collections.Counter(), collections.Counter()
for (ln, txt), (e, s) in per_line.items():
    f = fn_at[ln] if ln < len(fn_at) else "?"
    agg[f] += e
    smp[f] += s
for k, v in agg.most_common(30):
    print("%-26s %7.1f instr/thread  samples %d" % (k, v / W, smp[k]))
print()
for (ln, txt), (e, s) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:50]:
    print("%5d %6.1f %5d  %s" % (ln, e / W, s, txt))
