"""Summarise ncu artefacts (read here, no GPU needed) into small text files for profiles/.

    python tools/ncu_summary.py rep  gpurun_out/prof_X.ncu-rep  [n_envs]  > profiles/X.txt
    python tools/ncu_summary.py list gpurun_out/launches_X.csv            > profiles/launches_X.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def rep(path, n_envs=None):
    rows = ncu_csv(path, "raw")
    h, units, vals = rows[0], rows[1], rows[2]
    d = {k: (vals[i], units[i]) for i, k in enumerate(h)}
    print("# ncu --set full --clock-control none (cold cache, serialised replay) :", path.split("/")[-1])
    print("kernel:", d["Kernel Name"][0])
    for k in KEYS:
        if k in d:
            print("%-72s %s %s" % (k, d[k][0], d[k][1]))
    try:
        rd = float(d["dram__bytes_read.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][1]]
        print("dram traffic per launch (read + write)                                   %.1f MB" % ((rd + wr) / 1e6))
        if n_envs:
            print("algorithmic bytes per launch (349 B x %d envs)                      %.1f MB" % (n_envs, 349 * n_envs / 1e6))
            inst = float(d["smsp__inst_executed.sum"][0])
            print("warp-instructions per env-step thread                                    %.0f" % (inst * 32 / n_envs))
    except Exception as e:  # noqa: BLE001
        print("traffic: n/a", e)
    src = ncu_csv(path, "source", ["--print-source", "sass"])
    hdr, data = None, []
    for r in src:
        if r and r[0] == "Address":
            hdr = hdr or r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    if hdr:
        hi = {k: i for i, k in enumerate(hdr)}
        stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
        tot = collections.Counter()
        ops = collections.Counter()
        samples = 0
        for r in data:
            samples += int(r[hi["# Samples"]])
            for k in stalls:
                tot[k] += int(r[hi[k]])
            parts = r[hi["Source"]].strip().split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            ops[op.split(".")[0]] += int(r[hi["Instructions Executed"]])
        print("warp-state samples:", samples)
        for k, v in tot.most_common(10):
            if v:
                print("  %-26s %5.1f %%" % (k, 100.0 * v / max(samples, 1)))
        allops = sum(ops.values())
        print("SASS op mix (share of executed warp-instructions):")
        print("  " + "  ".join("%s %.1f%%" % (k, 100.0 * v / allops) for k, v in ops.most_common(18)))
        bulk = sum(v for k, v in ops.items() if k in ("UBLKCP", "SYNCS", "ACQBULK"))
        print("  TMA / mbarrier / PDL instructions executed (UBLKCP, SYNCS, ACQBULK): %d" % bulk)


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ik, iv, ig = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        d[(r[ik].split("(")[0][:90], r[ig])].append(float(r[iv].replace(",", "")))
    total = sum(sum(v) for v in d.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none :", path.split("/")[-1])
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print("%-92s %-14s %7s %10s %10s %7s" % ("kernel", "grid", "count", "avg ns", "min ns", "share"))
    for (k, g), v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("%-92s %-14s %7d %10.0f %10.0f %6.1f%%" % (k, g, len(v), sum(v) / len(v), min(v), 100.0 * sum(v) / total))


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        rep(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    else:
        launches(sys.argv[2])
