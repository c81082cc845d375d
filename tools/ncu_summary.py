#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key counters per kernel. usage: tools/ncu_summary.py file.ncu-rep [n_envs]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = float(sys.argv[2]) if len(sys.argv) > 2 else 1048576.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__sectors_read.sum', 'dram__sectors_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_lg.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'lts__t_sectors_srcunit_tex_lookup_miss.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_adu.sum', 'sm__inst_executed_pipe_uniform.sum']
for d in data:
    print('---', d[hdr.index('Kernel Name')][:70])
    for k in keys:
        if k in hdr:
            v = d[hdr.index(k)]
            try:
                per = float(v.replace(',', '')) / n
                extra = f'   per-env {per:10.3f}' if k.endswith('.sum') else ''
            except ValueError:
                extra = ''
            print(f'{k:85s} {v:>18s} {units[hdr.index(k)]:12s}{extra}')
