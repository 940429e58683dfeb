"""Summarise an ncu launch list (gpu__time_duration.sum CSV) of `bench.py --one-forward`: per-launch table of the LAST
forward in the file and a per-kernel aggregate.  python tools/launch_table.py gpurun_out/launches.csv [out.txt]"""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
def _us(x):      # gpu__time_duration.sum in the CSV's unit -> ns (the table below prints us)
    v = float(x['Metric Value'].replace(',', ''))
    return v * {'ns': 1.0, 'nsecond': 1.0, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6}.get(x.get('Metric Unit', 'ns'), 1.0)


# a CSV may carry several metrics per launch (tools/gpu_r2_profiles.sh adds the DRAM byte counters): keep the durations
rows = [(x['Kernel Name'], _us(x), x.get('Grid Size', ''))
        for x in csv.DictReader(lines) if x.get('Metric Name', 'gpu__time_duration.sum') == 'gpu__time_duration.sum']
starts = [i for i, x in enumerate(rows) if 'prep_burst' in x[0]]
fw = rows[starts[-1]:]
out = []
agg = collections.OrderedDict()
cum = 0.0
for i, (n, v, g) in enumerate(fw):
    cum += v
    short = n.split('(')[0].replace('void ', '').replace('dbsr::', '')[:44]
    out.append(f'{i:3d} {v / 1e3:8.1f} us  cum {cum / 1e3:9.1f}  grid {g:>14s}  {short}')
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += v
out.append('')
out.append(f'total {cum / 1e3:.1f} us over {len(fw)} launches (serialised, cold cache: compare shares, not absolutes)')
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f'{v / 1e3:9.1f} us  {100 * v / cum:5.1f} %  x{c:3d}  {k}')
text = '\n'.join(out)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(text + '\n')
print(text)
