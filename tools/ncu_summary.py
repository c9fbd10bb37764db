#!/usr/bin/env python
"""Condense ncu exports into the small CSVs kept under profiles/:

    python tools/ncu_summary.py full  gpurun_out/NAME.raw.csv [more.raw.csv ...] > profiles/rNN/ncu_full_kernels.summary.csv
    python tools/ncu_summary.py share gpurun_out/launches.csv                    > profiles/rNN/ncu_launch_share.csv

`full`: one row per captured kernel launch with the counters DESIGN.md quotes (from `ncu -i X.ncu-rep --page raw --csv`).
`share`: per-kernel launch count, mean duration and share of the summed GPU time (from the `--metrics
gpu__time_duration.sum` launch list)."""
import collections
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
]


def full(paths):
    w = csv.writer(sys.stdout)
    w.writerow(["source", "kernel", "metric", "unit", "value"])
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        for r in rows[2:]:
            name = r[ki].split("(")[0]
            for m in KEEP:
                if m in hdr:
                    i = hdr.index(m)
                    w.writerow([path.split("/")[-1], name, m, units[i], r[i]])


def share(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
        agg.setdefault(d["Kernel Name"].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "mean_us", "total_us", "share_pct"])
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        w.writerow([k, len(v), f"{sum(v) / len(v):.1f}", f"{sum(v):.1f}", f"{100 * sum(v) / tot:.2f}"])


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("full", "share"):
        sys.exit(__doc__)
    full(sys.argv[2:]) if sys.argv[1] == "full" else share(sys.argv[2])
