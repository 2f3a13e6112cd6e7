#!/bin/bash
# key counters of every kernel in an .ncu-rep (run here, no GPU needed)
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__cycles_elapsed.avg.per_second']
want+= [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h] + [h for h in hdr if h.startswith('smsp__average_warp') and 'pct' not in h and 'issue_stalled' in h and 'per_warp_active' in h]
idx={h:i for i,h in enumerate(hdr)}
for r in rows[2:]:
    print('==', r[idx['Kernel Name']].split('(')[0], ' (id', r[idx['ID']] if 'ID' in idx else '', ')')
    for w in want[1:]:
        if w in idx:
            v=r[idx[w]]
            try:
                fv=float(v.replace(',',''))
                if 'issue_stalled' in w and fv<0.15: continue
            except: pass
            print('   %-95s %s %s'%(w, v, units[idx[w]]))
"
