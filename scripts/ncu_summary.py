"""Condensed text summary of an .ncu-rep (first captured launch): the numbers DESIGN.md / bench cite."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, v = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma', 'sm__pipe_fma_cycles_active', 'sm__pipe_fmaheavy', 'sm__pipe_fmalite',
        'sm__inst_executed_pipe_fp64', 'sm__inst_executed_pipe_lsu',
        'sm__inst_executed_pipe_alu', 'sm__inst_executed_pipe_xu',
        'sm__pipe_tensor_cycles_active', 'sm__inst_executed_pipe_tensor',
        'smsp__issue_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__average_warp', 'smsp__average_warps_issue_stalled', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__pcsamp_warps_issue_stalled']
for i, name in enumerate(h):
    if any(name.startswith(w) for w in want):
        if 'dshared' in name or 'per_second' in name: continue
        print(f'{name} = {v[i]} {units[i]}')
