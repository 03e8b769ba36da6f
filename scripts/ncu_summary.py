"""Extract the judged metrics from an .ncu-rep (ncu --set full) into a small text table."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed.sum', 'sm__cycles_active.avg',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64_op_dmma.sum',
        'smsp__inst_executed_pipe_fp64.sum']
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print(f'# {rep}: {len(rows) - 2} launch(es) of', sorted({r[hdr.index("Kernel Name")][:80] for r in rows[2:]}))
extra = [h for h in hdr if ('pipe_tensor' in h or 'dmma' in h.lower() or 'pipe_fp64' in h) and h not in KEYS
         and (h.endswith('.sum') or h.endswith('pct_of_peak_sustained_active'))][:24]
for k in KEYS + extra:
    if k in hdr:
        i = hdr.index(k)
        print(f'{k:90s} {units[i]:14s} {[r[i] for r in rows[2:]]}')
