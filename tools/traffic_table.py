"""Aggregate an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch (one eager
forward, tools/gpu_ncu_traffic.sh) into per-kernel-family DRAM traffic.  python tools/traffic_table.py in.csv out.json"""
import collections
import csv
import json
import sys

FAMILY = [('resblock32_tc', 'resblock_tc'), ('conv_tc', 'conv_tc'), ('softmax_wsum', 'softmax_wsum'), ('corr81', 'corr81'), ('blur3x3', 'blur3x3'),
          ('warp_proj', 'warp_proj'), ('predictor', 'predictor'), ('space_to_depth', 'copy'), ('copy_channels', 'copy'),
          ('deconv', 'deconv'), ('flow_from_taps', 'deconv'), ('prep_burst', 'prep_burst'), ('offsets_mod', 'offsets_mod'), ('flow_head', 'flow_head'),
          ('conv_direct', 'conv_direct')]


def unit_scale(u):
    return {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0,
            'nsecond': 1e-3, 'msecond': 1e3}[u]


with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
launch = collections.OrderedDict()
for r in csv.DictReader(lines):
    key = int(r['ID'])
    d = launch.setdefault(key, {'name': r['Kernel Name']})
    d[r['Metric Name']] = float(r['Metric Value'].replace(',', '')) * unit_scale(r['Metric Unit'])
ids = list(launch)
starts = [i for i in ids if 'prep_burst' in launch[i]['name']]
fw = [launch[i] for i in ids if i >= starts[-1]]
fam = collections.OrderedDict()
for d in fw:
    name = next((f for k, f in FAMILY if k in d['name']), 'other')
    a = fam.setdefault(name, {'launches': 0, 'dram_read_bytes': 0.0, 'dram_write_bytes': 0.0, 'us': 0.0})
    a['launches'] += 1
    a['dram_read_bytes'] += d.get('dram__bytes_read.sum', 0.0)
    a['dram_write_bytes'] += d.get('dram__bytes_write.sum', 0.0)
    a['us'] += d.get('gpu__time_duration.sum', 0.0)
out = {'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum, one eager forward, '
                 'serialised launches with cold caches', 'launches': len(fw), 'families': fam}
json.dump(out, open(sys.argv[2], 'w'), indent=1)
for k, a in fam.items():
    print(f"{k:14s} x{a['launches']:3d} {a['us']:9.1f} us  read {a['dram_read_bytes'] / 1e6:9.1f} MB  write {a['dram_write_bytes'] / 1e6:9.1f} MB")
