"""Extract the judged metrics of an .ncu-rep into a small text summary (kept under profiles/).
   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.summary.txt"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_requests_srcunit_tex_op_red.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "dram__bytes.sum.per_second", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_red.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "smsp__sass_inst_executed_op_tma_red.sum",
        "smsp__inst_executed_op_global_red.sum", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
        "sass__inst_executed_global_loads", "sass__inst_executed_global_stores", "launch__shared_mem_per_block_dynamic",
        "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print(f"{k:95s} {d[k]:>18s} {units[hdr.index(k)]}")
