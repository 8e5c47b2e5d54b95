"""Excerpt of an `ncu --set full` report: per captured launch the metrics DESIGN.md argues with.
usage: ncu -i X.ncu-rep --page raw --csv > X.csv; python profiles/summarize_full.py X.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_ldgsts.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("--")
    print(f"  {'Kernel Name':70s} {d.get('Kernel Name', '')[:90]}")
    for k in WANT:
        if k in d:
            print(f"  {k:70s} {d[k]} {u.get(k, '')}")
    st = []
    for k, v in d.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    tot = sum(x for x, _ in st) or 1.0
    st.sort(reverse=True)
    print("  stall reasons (share of warp-cycles per issue): " + ", ".join(f"{n} {100 * x / tot:.0f}%" for x, n in st[:7]))
