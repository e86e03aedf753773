"""Key figures of every kernel in an .ncu-rep (raw page): time, DRAM bytes, issue/occupancy, top stall reasons."""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default="-"):
    i = col.get(name)
    return r[i] if i is not None and r[i] != "" else default
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ_regs"),
        ("launch__occupancy_limit_shared_mem", "occ_smem"), ("launch__occupancy_limit_warps", "occ_warps"),
        ("smsp__inst_executed.sum", "inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "st_sleep"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar")]
for r in data:
    print("==", g(r, "Kernel Name")[:100], g(r, "Grid Size"), g(r, "Block Size"))
    print("   " + "  ".join(f"{lab}={g(r, m)}{'' if col.get(m) is None else units[col[m]]}" for m, lab in want))
