"""Print the headline metrics of every kernel in an .ncu-rep (ncu --set full capture): python tools/ncu_summary.py FILE"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_shared_ld.sum"]
for r in rows[2:]:
    print("KERNEL", r[hdr.index("Kernel Name")][:80])
    for i, h in enumerate(hdr):
        if h in want or "issue_stalled" in h and h.endswith("per_issue_active.ratio") or "bank_conflict" in h:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v != 0:
                print(f"   {h} [{units[i]}] = {r[i]}")
