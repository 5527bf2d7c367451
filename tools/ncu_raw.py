"""Key metrics of an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv | python tools/ncu_raw.py [extra-regex]`."""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
KEYS = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_blocks', 'launch__occupancy_limit_warps', 'launch__occupancy_limit_shared_mem',
        'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum']
for r in rows[2:]:
    print('-----')
    for i, h in enumerate(hdr):
        if h in KEYS or (pat and pat.search(h)):
            print(f"{h:72s} {r[i][:90]} {units[i]}")
