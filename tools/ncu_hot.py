"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv --print-source sass): totals per stall
reason and the hottest SASS instructions."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# kernels are concatenated: header line starts with "Address"
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
b = blocks[0]
h = {k: i for i, k in enumerate(b["hdr"])}
stalls = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
tot = {k: 0 for k in stalls}
samples = 0
execd = 0
for r in b["rows"]:
    samples += int(r[h["# Samples"]])
    execd += int(r[h["Instructions Executed"]])
    for k in stalls:
        tot[k] += int(r[h[k]])
print(b["name"])
print("samples", samples, "warp-instructions executed", execd)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-24s %7d  %5.1f%%" % (k, v, 100.0 * v / samples))
print("hottest instructions:")
order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][h["# Samples"]]))[:top]
for i in sorted(order):
    r = b["rows"][i]
    main = max(stalls, key=lambda k: int(r[h[k]]))
    print("  %5d %6s  %-22s %s" % (i, r[h["# Samples"]], main, r[h["Source"]].strip()[:90]))
