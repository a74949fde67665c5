"""Print the key --set full metrics (and the stall breakdown) of every kernel in an `ncu --page raw --csv` dump."""
import csv
import sys

rr = list(csv.reader(open(sys.argv[1])))
h, u = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "lts__t_sector_hit_rate.pct"]
for r in rr[2:]:
    print("==", r[h.index("Kernel Name")][:90])
    for w in want:
        if w in h:
            print("  %-70s %s %s" % (w, r[h.index(w)], u[h.index(w)]))
    st = []
    for i, x in enumerate(h):
        if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio"):
            try:
                st.append((float(r[i].replace(",", "")), x[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("  stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
