"""Summarise ncu outputs brought back in gpurun_out/ into small committed text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.txt
  python tools/ncu_summary.py full     gpurun_out/prof_X.ncu-rep  profiles/X_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        key = re.sub(r"\(.*", "", row["Kernel Name"])[:70] + " grid=" + row.get("Grid Size", "")
        agg.setdefault(key, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({src})\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:95s} n={len(v):4d} mean_us={sum(v)/len(v):9.2f} total_ms={sum(v)/1000:8.3f} share={sum(v)/tot:6.3f}\n")
        f.write(f"total_ms={tot/1000:.3f}\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# selected metrics from `ncu --set full --clock-control none` ({src}); one block per profiled launch\n")
        for r in rows[2:]:
            f.write("---\n" + r[hdr.index("Kernel Name")] + "\n")
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k:80s} {r[i]:>16s} {units[i]}\n")
    print(open(dst).read())


def _to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def traffic(src, dst, note=""):
    """profiles/ncu_traffic.json: DRAM bytes per launch of every kernel in an `ncu --set full` capture (mean over the
    captured launches) — bench.py copies the dominant kernel's figure into `roofline.traffic`."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, it, ik = (hdr.index(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "Kernel Name"))
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ik]).split("::")[-1].split("<")[0].replace("void ", "").strip()
        agg.setdefault(name, []).append((_to_bytes(r[ir], units[ir]), _to_bytes(r[iw], units[iw]), float(r[it].replace(",", "")), units[it]))
    db = json.load(open(dst)) if os.path.exists(dst) else {}
    for name, v in agg.items():
        n = len(v)
        tm = sum(x[2] for x in v) / n
        tm_ms = tm if v[0][3] == "ms" else tm / 1e3 if v[0][3] == "us" else tm / 1e6
        db[name] = {"dram_bytes_read_per_launch": int(sum(x[0] for x in v) / n), "dram_bytes_write_per_launch": int(sum(x[1] for x in v) / n),
                    "dram_bytes_per_launch": int(sum(x[0] + x[1] for x in v) / n), "launches_captured": n, "ncu_gpu_time_ms": round(tm_ms, 5),
                    "source": os.path.basename(src), "note": note}
    json.dump(db, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps({k: db[k] for k in agg}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
