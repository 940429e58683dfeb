"""Print the metrics of an .ncu-rep that the profile summaries quote (read here, no GPU): python tools/ncu_summary.py rep [rep ...]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
        'smsp__cycles_active.avg', 'sm__sass_inst_executed_op_shared_ld.sum', 'sm__sass_inst_executed_op_shared_st.sum',
        'smsp__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_uniform.sum', 'gpc__cycles_elapsed.avg.per_second',
        'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active_mem_lg.sum']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print('==', rep)
    for i, h in enumerate(hdr):
        if h == 'Kernel Name':
            print('  kernel:', vals[i][:110])
        if h in WANT or ('stall' in h and 'per_warp_active' in h and float(vals[i] or 0) >= 3.0):
            print(f'  {h:75s} {vals[i]:>16s} {units[i]}')
