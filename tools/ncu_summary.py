"""Key metrics of every kernel in an ncu report.  Usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
        'sm__cycles_elapsed.avg', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U = rows[0], rows[1]
ki = H.index('Kernel Name')
for r in rows[2:]:
    print('==', r[ki][:110])
    for i, h in enumerate(H):
        if h in WANT or (len(sys.argv) > 2 and sys.argv[2] in h):
            print('   %-85s %s %s' % (h, r[i], U[i]))
