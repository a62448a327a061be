"""Selected `ncu --set full` metrics of every kernel in a report as one CSV (metric rows x kernel columns), the form of
profiles/r0N_ncu_*_metrics.csv:   python scripts/ncu_metrics_csv.py gpurun_out/x.ncu-rep > profiles/r02_ncu_metrics.csv"""
import csv
import io
import re
import subprocess
import sys

METRICS = """launch__grid_size launch__block_size launch__registers_per_thread launch__shared_mem_per_block_dynamic
launch__shared_mem_per_block_static launch__occupancy_limit_registers launch__occupancy_limit_shared_mem gpu__time_duration.sum
sm__cycles_active.avg sm__cycles_elapsed.max sm__cycles_elapsed.avg.per_second sm__warps_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
sm__ops_path_tensor_src_fp64.sum smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__warps_eligible.avg.per_cycle_active smsp__warps_active.avg.per_cycle_active dram__bytes_read.sum dram__bytes_write.sum
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio""".split()

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, launches = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
names = [re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void <unnamed>::", "") for r in launches]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + names)
for m in METRICS:
    if m in col:
        w.writerow([m, units[col[m]]] + [r[col[m]] for r in launches])
