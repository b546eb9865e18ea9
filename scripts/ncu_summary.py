"""Summarise an `ncu --page raw --csv` export: one block per kernel launch with the metrics the roofline uses."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max']
stall = [h for h in hdr if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
for r in data:
    print('##', r[idx['Kernel Name']][:100])
    for w in want:
        if w in idx:
            print(f'   {w:72s} {r[idx[w]]:>16s} {units[idx[w]]}')
    top = sorted(((float(r[idx[n]]), n) for n in stall), reverse=True)[:6]
    print('   top stalls (warps per issue): ' + ', '.join(f"{n.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, n in top))
