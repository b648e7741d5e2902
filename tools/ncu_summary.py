#!/usr/bin/env python
"""Summarises ncu artefacts brought back in gpurun_out/ into small text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_ck2_r01.csv  > profiles/r01_ck2_launches.txt
    python tools/ncu_summary.py full     gpurun_out/prof_ck2_r01.ncu-rep  > profiles/r01_ck2_ncu_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        key = (row["Kernel Name"], row["Grid Size"], row["Block Size"])
        agg.setdefault(key, []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare shares)")
    print("# source: %s ; total %.1f us over %d launches" % (path, tot / 1e3, sum(len(v) for v in agg.values())))
    print("%-90s %-14s %-12s %6s %12s %10s %7s" % ("kernel", "grid", "block", "n", "sum_us", "avg_us", "share"))
    for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-90s %-14s %-12s %6d %12.1f %10.2f %7.3f" % (name[:90], grid, block, len(v), sum(v) / 1e3,
                                                           sum(v) / len(v) / 1e3, sum(v) / tot))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none --import-source on ; source: %s" % path)
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("## launch id %s : %s" % (r[0], r[name_i]))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-75s %-16s %s" % (k, units[i], "  ".join(r[i] for r in rows[2:])))
    det = subprocess.run(["ncu", "-i", path, "--page", "details"], capture_output=True, text=True).stdout
    print("\n# --page details of the first captured launch (sections: SOL, compute, memory, scheduler, warp state, occupancy)")
    seen = 0
    for line in det.splitlines():
        if "Context 1, Stream" in line:
            seen += 1
        if seen > 1:
            break
        if line.strip():
            print(line.rstrip()[:160])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
