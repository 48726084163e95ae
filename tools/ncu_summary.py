"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py file.ncu-rep [extra-metric-regex]"""
import csv, re, subprocess, sys, io
rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for v in vals:
    print("==", v[hdr.index("Kernel Name")][:110], "| id", v[hdr.index("ID")])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k); print(f"   {k:75s} {v[i]:>16s} {units[i]}")
    st = [(float(v[i]), h) for i, h in enumerate(hdr) if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", h) and v[i]]
    for x, h in sorted(st, reverse=True)[:7]:
        print(f"   stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {x:8.2f}")
    if extra:
        for i, h in enumerate(hdr):
            if re.search(extra, h) and v[i] not in ("", "0", "0.000000"):
                print(f"   {h:90s} {v[i]:>16s} {units[i]}")
