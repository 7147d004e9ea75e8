import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
idx = {h: i for i, h in enumerate(rows[0])}
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[idx['ID']], {'name': r[idx['Kernel Name']][:34]})[r[idx['Metric Name']]] = (r[idx['Metric Value']], r[idx['Metric Unit']])
def f(v, m):
    x, u = v.get(m, ('0', ''))
    x = float(x.replace(',', ''))
    return x * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
for k, v in d.items():
    print(k, v['name'], 'grid', v['launch__grid_size'][0], 'smem', v['launch__shared_mem_per_block_dynamic'][0], v['launch__shared_mem_per_block_dynamic'][1], 'regs', v['launch__registers_per_thread'][0],
          'ms %.3f' % f(v, 'gpu__time_duration.sum'), 'Minst %.0f' % (f(v, 'smsp__inst_executed.sum') / 1e6),
          'issue%% %.1f' % f(v, 'smsp__issue_active.avg.pct_of_peak_sustained_active'), 'warps/sched %.2f' % f(v, 'smsp__warps_active.avg.per_cycle_active'),
          'dramR %.0f MB dramW %.0f MB' % (f(v, 'dram__bytes_read.sum'), f(v, 'dram__bytes_write.sum')), 'L2hit %.0f L1hit %.0f' % (f(v, 'lts__t_sector_hit_rate.pct'), f(v, 'l1tex__t_sector_hit_rate.pct')))
