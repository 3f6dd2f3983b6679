"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one markdown table row per captured launch with the metrics the
roofline discussion uses, and (optionally) the DRAM-traffic entries bench.py quotes (profiles/r02_traffic.json).
usage: python scripts/ncu_summary.py raw.csv [--traffic-key C4/spmv_dot/n1,C4/spmv_tdot/n1 --traffic-json profiles/r02_traffic.json]"""
import csv, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}
def get(r, name, default=float('nan')):
    if name not in hdr:
        return default
    i = hdr.index(name)
    try:
        return float(r[i].replace(',', '')) * scale.get(units[i], 1.0)
    except ValueError:
        return default
want = [('time us', 'gpu__time_duration.sum'), ('DRAM read MB', 'dram__bytes_read.sum'), ('DRAM write MB', 'dram__bytes_write.sum'),
        ('DRAM % of peak', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        ('LSU data pipe % busy', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
        ('shared wavefronts', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
        ('shared bank conflicts', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
        ('warp instructions', 'smsp__inst_executed.sum'), ('issue active %', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        ('registers/thread', 'launch__registers_per_thread'), ('tensor pipe % active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('fp64 pipe % active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')]
kname = hdr.index('Kernel Name')
print('| launch | kernel | ' + ' | '.join(w[0] for w in want) + ' |')
print('|---|---|' + '---|' * len(want))
out = []
for n, r in enumerate(body):
    vals = []
    for label, m in want:
        v = get(r, m)
        if label.endswith('MB'):
            v /= 1e6
        vals.append(v)
    out.append(vals)
    print('| %d | `%s` | ' % (n, r[kname].split('(')[0][:40]) + ' | '.join(('%.4g' % v) if v == v else '-' for v in vals) + ' |')
if '--traffic-key' in sys.argv:
    import bench
    keys = sys.argv[sys.argv.index('--traffic-key') + 1].split(',')
    path = sys.argv[sys.argv.index('--traffic-json') + 1]
    tr = json.load(open(path)) if os.path.exists(path) else {}
    for key, r in zip(keys, body):
        tr[key] = {'dram_bytes': get(r, 'dram__bytes_read.sum') + get(r, 'dram__bytes_write.sum'),
                   'time_us_under_ncu': get(r, 'gpu__time_duration.sum'),
                   'kernel_sources': list(bench.KERNEL_SOURCES[key.split('/')[1]]),
                   'kernel_version': bench.kernel_version(bench.KERNEL_SOURCES[key.split('/')[1]]),
                   'source': '%s (ncu --set full --clock-control none, launch of %s)' % (os.path.basename(sys.argv[1]).replace('.csv', '.ncu-rep'), r[kname].split('(')[0])}
    json.dump(tr, open(path, 'w'), indent=1)
    print('traffic entries written to', path)
