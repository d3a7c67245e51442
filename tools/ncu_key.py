"""Key counters of one `ncu --set full` capture: python tools/ncu_key.py gpurun_out/prof_x.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, val = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, val)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum.per_cycle_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg']
for k in keys:
    for h in d:
        if h == k or h.startswith(k + ' '):
            print(f'{h:80s} {d[h][0]:>16s} {d[h][1]}')
print('--- stalls per issue')
st = [(float(v[0]), h) for h, v in d.items() if 'issue_stalled' in h and 'per_issue_active' in h and v[0]]
for v, h in sorted(st, reverse=True)[:10]:
    print(f'   {h.split("issue_stalled_")[1].split("_per_issue")[0]:24s} {v:.3f}')
