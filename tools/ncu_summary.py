"""Condense an .ncu-rep into the small CSV kept under profiles/ (metric, unit, value per kernel launch).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x_ncu.csv
"""
import csv
import re
import subprocess
import sys

KEEP = [
    r"^dram__bytes_(read|write)\.sum$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^gpu__time_duration\.sum$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$",
    r"^l1tex__t_(sectors|requests)_pipe_lsu_mem_(global|local)_op_(ld|st)\.sum$", r"^lts__t_sector_hit_rate\.pct$",
    r"^lts__t_sectors_srcunit_tex_op_read\.sum$", r"^launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic)$",
    r"^launch__occupancy_limit_(registers|shared_mem)$", r"^sm__cycles_elapsed\.avg$", r"^sm__inst_executed\.avg\.per_cycle_elapsed$",
    r"^sm__inst_executed_pipe_(alu|fma|lsu|tc|tmem|uniform|xu)\.avg\.pct_of_peak_sustained_active$",
    r"^sm__pipe_(alu|fma|tensor)_cycles_active\.avg\.pct_of_peak_sustained_active$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
    r"^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        for vals in rows[2:]:
            w.writerow(["Kernel Name", "", vals[name_col]])
            for h, u, v in zip(hdr, units, vals):
                if any(re.search(k, h) for k in KEEP):
                    w.writerow([h, u, v])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
