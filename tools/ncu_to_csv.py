"""Summarise an `ncu --set full` capture (.ncu-rep) as the metric,unit,value CSV kept under profiles/:
    python tools/ncu_to_csv.py gpurun_out/X.ncu-rep profiles/rNN_name_ncu.csv
The metric list is the one of the round-1 summaries plus the issue / stall figures the design notes quote."""
import csv, subprocess, sys

KEEP_PREFIX = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
               "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
               "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
               "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
               "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
               "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "launch__block_size", "launch__grid_size",
               "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
               "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
               "sm__cycles_elapsed.avg", "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.avg.per_cycle_active",
               "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
               "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
               "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
               "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_local_ld.sum",
               "smsp__sass_inst_executed_op_local_st.sum")


def main(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows[2:]:
            w.writerow(["Kernel Name", "", r[hdr.index("Kernel Name")]])
            for i, h in enumerate(hdr):
                if h in KEEP_PREFIX or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                    w.writerow([h, units[i], r[i]])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
