#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/NAME.txt
  python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  profiles/NAME.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        d[name][0] += 1
        d[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in d.values())
    with open(dst, "w") as f:
        f.write("# per-kernel totals from `ncu --metrics gpu__time_duration.sum --clock-control none` (%s)\n" % src)
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        f.write("%-60s %8s %12s %8s %12s\n" % ("kernel", "launches", "total_ms", "share", "avg_us"))
        for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.3f %7.1f%% %12.2f\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e3))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# selected metrics from `ncu --set full --clock-control none` (%s)\n" % src)
        for r in rows[2:]:
            f.write("\n== %s  grid=%s block=%s\n" % (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "?",
                                                  r[hdr.index("Block Size")] if "Block Size" in hdr else "?"))
            for m in KEEP:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("  %-72s %16s %s\n" % (m, r[i], units[i]))
            if "dram__bytes_read.sum" in hdr:
                def val(name):
                    i = hdr.index(name)
                    v = float(r[i].replace(",", ""))
                    u = units[i].lower()
                    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1, "tbyte": 1e12}.get(u, 1)
                tr = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
                ti = hdr.index("gpu__time_duration.sum")
                t = float(r[ti].replace(",", "")) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(units[ti], 1e-9)
                f.write("  %-72s %16.0f byte\n" % ("traffic = dram read + write (per launch)", tr))
                f.write("  %-72s %16.1f GB/s (under ncu replay clocks)\n" % ("dram traffic / duration", tr / t / 1e9))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
