"""Per-opcode executed-instruction counts per thread from an ncu source-page CSV."""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
warps = float(sys.argv[2]) if len(sys.argv) > 2 else 131072.0
hdr, data = None, []
for r in rows:
    if r and r[0] == "Address":
        if hdr is None:
            hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
h = {k: i for i, k in enumerate(hdr)}
ops = collections.Counter()
tot = 0
for r in data:
    n = int(r[h["Instructions Executed"]])
    parts = r[h["Source"]].strip().split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    ops[op.split(".")[0]] += n
    tot += n
print("per-thread instructions %.0f" % (tot / warps))
print("  ".join("%s %.0f" % (k, v / warps) for k, v in ops.most_common(45)))
