"""Writes profiles/r01_traffic.json (dram bytes per launch of the persistent generation kernel at the bench workload) from an
ncu CSV produced by:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:wn_persistent --csv \
      --log-file gpurun_out/traffic.csv python scripts/ncu_target.py 48000
(two metrics = a single replay pass, so the 1.6 s launch is not replayed 40 times)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'traffic.csv')
rows = [r for r in csv.DictReader(l for l in open(src) if l.startswith('"'))]
per = {}
for r in rows:
    per.setdefault(r['ID'], {})[r['Metric Name']] = (float(r['Metric Value']), r['Metric Unit'])
unit = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
launches = []
for k, m in per.items():
    tot = sum(v * unit[u] for v, u in m.values())
    launches.append(tot)
out = {"kernel": rows[0]['Kernel Name'].split('(')[0], "launches_captured": len(launches), "dram_bytes_per_launch": launches[-1],
       "steps_per_launch": 48000, "rows": 8, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, last captured launch"}
json.dump(out, open(os.path.join(ROOT, 'profiles', 'r01_traffic.json'), 'w'), indent=1)
print(out)
