"""Summarise an ncu launch list (csv, long format: one row per (launch, metric)) per kernel name.
usage: summarize_launches.py gpurun_out/launches_X.csv [--pass-launches N]"""
import collections
import csv
import sys

path = sys.argv[1]
rows = list(csv.DictReader(l for l in open(path) if not l.startswith('==')))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(int(r['ID']), {'name': r['Kernel Name'], 'grid': r['Grid Size'], 'block': r['Block Size']})
    try:
        v = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        v = float('nan')
    unit = r['Metric Unit']
    name = r['Metric Name']
    if name == 'gpu__time_duration.sum':
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1.0)
    if name.startswith('dram__bytes'):
        v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
    d[name] = v
def short(n):
    n = n.replace('tod::', '')
    return n.split('(')[0][:60]
agg = collections.OrderedDict()
for i, d in launch.items():
    a = agg.setdefault(short(d['name']), {'n': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0, 'tensor': 0.0})
    a['n'] += 1
    a['us'] += d.get('gpu__time_duration.sum', 0.0)
    a['rd'] += d.get('dram__bytes_read.sum', 0.0)
    a['wr'] += d.get('dram__bytes_write.sum', 0.0)
    t = d.get('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed')
    if t == t and t is not None:
        a['tensor'] += t * d.get('gpu__time_duration.sum', 0.0)
tot = sum(a['us'] for a in agg.values())
print(f"{len(launch)} launches, {tot:.1f} us summed (cold-cache, serialised: compare shares, not absolutes)")
print(f"{'kernel':60s} {'launches':>8s} {'sum us':>10s} {'share':>7s} {'avg us':>8s} {'DRAM rd MB/launch':>18s} {'wr MB/launch':>13s} {'GB/s':>8s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    bw = (a['rd'] + a['wr']) / max(a['us'], 1e-9) / 1e3
    print(f"{k:60s} {a['n']:8d} {a['us']:10.1f} {100 * a['us'] / tot:6.1f}% {a['us'] / a['n']:8.1f} {a['rd'] / a['n'] / 1e6:18.2f} {a['wr'] / a['n'] / 1e6:13.2f} {bw:8.0f}")
