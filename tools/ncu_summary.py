#!/usr/bin/env python
"""Turn ncu output into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches_summary.csv
      per-kernel launch count / total / share / average from an
      `ncu --metrics gpu__time_duration.sum --csv --log-file ...` launch list
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/rNN_ncu_full_metrics.txt
      the handful of `--set full` counters the roofline argument uses, per captured launch
"""
import collections
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[row["Metric Unit"]]
        a = agg.setdefault(row["Kernel Name"].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print("kernel,launches,total_us,share_pct,avg_us")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'"{k}",{n},{t:.1f},{100 * t / tot:.1f},{t / n:.1f}')


def shares(path, workload="ppo_ant_4096x32"):
    """profiles/rNN_shares.json: what bench.py embeds as `shares` (kernel -> share of the summed kernel time of the launch list)."""
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[row["Metric Unit"]]
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    out = {"workload": workload,
           "source": "ncu --metrics gpu__time_duration.sum --clock-control none launch list of `python bench.py --steps 2 --warmup 3 "
                     "--no-cpu-baseline --no-kernel-sweep --no-recipes` (cold-cache, serialised launches: shares, not absolute times)",
           "total_us": round(tot, 1),
           "kernels": {k: {"launches": n, "total_us": round(t, 1), "share": round(t / tot, 4), "avg_us": round(t / n, 2)}
                       for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]) if t / tot >= 0.001}}
    print(json.dumps(out, indent=1))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        grid = r[idx["Grid Size"]] if "Grid Size" in idx else ""
        print(f"--- {r[idx['Kernel Name']]} grid {grid}")
        for m in FULL_METRICS:
            if m in idx:
                print(f"  {m:78s} {r[idx[m]]:>18s} {units[idx[m]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full, "shares": shares}[sys.argv[1]](*sys.argv[2:])
