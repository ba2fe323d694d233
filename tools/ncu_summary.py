#!/usr/bin/env python3
"""Summarise ncu output into profiles/ (run here, no GPU needed).
    python tools/ncu_summary.py launches gpurun_out/r1_launches_v5.csv "<comment>" > profiles/x.csv
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep "<comment>" > profiles/x.txt"""
import csv
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg")


def launches(path, comment):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, order = {}, []
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        if name not in agg:
            agg[name] = [0, 0.0]
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# " + comment)
    print("# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised "
          "under ncu: compare SHARES. unit: ns")
    print("kernel,launches,total_ns,share")
    for name in sorted(agg, key=lambda n: -agg[n][1]):
        print("%s,%d,%d,%.4f" % (name, agg[name][0], agg[name][1], agg[name][1] / tot))


def full(path, comment):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("# " + comment)
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name" or any(h.startswith(k) for k in KEEP):
            print("%s [%s] = %s" % (h, u, v))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
