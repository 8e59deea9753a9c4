"""Per-kernel roofline table from an ncu CSV (run here, no GPU needed):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed \
        --clock-control none --csv --log-file X.csv python tools/step_once.py c2
    python tools/roofline_table.py X.csv [hbm_peak_gbs=6545.6] [skip_first_n_launches]
Times are ncu's (cold caches, serialised launches): use the shares and the per-launch fractions, not the sums."""
import collections, csv, sys
path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6545.6
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lines = [l for l in open(path) if not l.startswith('==')]
per = collections.OrderedDict()
for row in csv.DictReader(lines):
    kid = int(row['ID'])
    d = per.setdefault(kid, {'name': row['Kernel Name']})
    v = float(row['Metric Value'].replace(',', '') or 0)
    u = row['Metric Unit']
    m = row['Metric Name']
    if m == 'gpu__time_duration.sum':
        v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
        d['us'] = v
    elif m.startswith('dram__bytes'):
        mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        d[m] = v * mul
    else:
        d[m] = v
agg = collections.OrderedDict()
for kid, d in per.items():
    if kid < skip or 'us' not in d:
        continue
    n = d['name']
    n = n[n.find('::', n.find('unnamed')) + 2:] if 'unnamed' in n else n
    n = n.split('(')[0][:44]
    a = agg.setdefault(n, {'n': 0, 'us': 0.0, 'bytes': 0.0, 'tensor': 0.0, 'sm': 0.0})
    a['n'] += 1
    a['us'] += d['us']
    a['bytes'] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    a['tensor'] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0) * d['us']
    a['sm'] += d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0) * d['us']
tot = sum(a['us'] for a in agg.values())
print(f'{"kernel":44s} {"n":>4s} {"us/launch":>10s} {"share":>6s} {"DRAM MB/launch":>14s} {"GB/s":>8s} {"%HBM":>6s} {"tensor%":>8s} {"SM%":>6s}')
for n, a in sorted(agg.items(), key=lambda x: -x[1]['us']):
    us = a['us'] / a['n']
    mb = a['bytes'] / a['n'] / 1e6
    gbs = a['bytes'] / (a['us'] * 1e-6) / 1e9 if a['us'] else 0
    print(f'{n:44s} {a["n"]:4d} {us:10.1f} {100 * a["us"] / tot:5.1f}% {mb:14.2f} {gbs:8.0f} {100 * gbs / peak:5.1f}% {a["tensor"] / a["us"]:8.1f} {a["sm"] / a["us"]:6.1f}')
