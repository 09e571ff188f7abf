#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py gpurun_out/x.ncu-rep [more metrics regex]"""
import csv, subprocess, sys, re, io
rep = sys.argv[1]; extra = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
pat = re.compile(r'^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|gpu__dram_throughput\.avg\.pct|launch__registers_per_thread|launch__grid_size|launch__block_size|launch__occupancy_limit|launch__waves|sm__warps_active\.avg\.pct|sm__throughput\.avg\.pct|sm__inst_executed_pipe_fp64|sm__pipe_fp64_cycles_active\.avg\.pct|smsp__inst_executed\.sum$|smsp__inst_executed_pipe_(lsu|fp64|alu|fma|xu)\.sum$|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate\.pct|smsp__issue_active\.avg\.pct|l1tex__throughput\.avg\.pct|lts__throughput\.avg\.pct|smsp__average_warps?_issue_stalled.*_per_issue_active|smsp__average_warp_latency_issue_stalled|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|smsp__thread_inst_executed_per_inst_executed\.ratio|sm__cycles_active\.avg$|l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum$|l1tex__t_requests_pipe_lsu_mem_global_op_ld\.sum$)' + ('|' + extra if extra else ''))
for r in rows[2:]:
    for i, name in enumerate(h):
        if pat.search(name):
            v = r[i]
            if 'stalled' in name:
                try:
                    if float(v.replace(',', '')) < 0.15: continue
                except ValueError: pass
            print('  %-95s %18s %s' % (name, v, rows[1][i]))
    print('-' * 60)
