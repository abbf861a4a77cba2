"""Print the roofline-relevant metrics of every kernel in an `ncu -i X.ncu-rep --page raw --csv` dump."""
import csv
import re
import sys

KEYS = [r'^gpu__time_duration\.sum$', r'^dram__bytes_read\.sum$', r'^dram__bytes_write\.sum$',
        r'^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$', r'^lts__t_sector_hit_rate\.pct$',
        r'^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$',
        r'^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed$',
        r'^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active$',
        r'^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$', r'^smsp__issue_active\.avg\.pct_of_peak_sustained_active$',
        r'^sm__warps_active\.avg\.pct_of_peak_sustained_active$', r'^launch__registers_per_thread$',
        r'^launch__grid_size$', r'^launch__block_size$', r'^launch__shared_mem_per_block_dynamic$',
        r'^l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum$', r'^l1tex__t_requests_pipe_lsu_mem_global_op_ld\.sum$',
        r'^l1tex__t_sectors_pipe_lsu_mem_global_op_st\.sum$', r'^l1tex__t_requests_pipe_lsu_mem_global_op_st\.sum$',
        r'^sm__cycles_active\.avg$', r'^smsp__inst_executed\.sum$']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d.get('Kernel Name', '?')[:90])
    for k in hdr:
        if any(re.search(p, k) for p in KEYS) and d[k] not in ('', 'n/a'):
            print(f'   {k:75s} {d[k]:>16s} {units[hdr.index(k)]}')
