#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): key memory/occupancy/stall metrics per kernel."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max', 'launch__waves_per_multiprocessor',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')]
        print(f'== {rep}: {name[:70]}')
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f'   {h:75s} {vals[i]:>18s} {units[i]}')
        st = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
                try:
                    v = float(vals[i])
                except ValueError:
                    continue
                if v > 0.25:
                    st.append((v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        print('   stalls/issue:', ', '.join(f'{n} {v:.2f}' for v, n in sorted(st, reverse=True)))
