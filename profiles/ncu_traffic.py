#!/usr/bin/env python
"""Write profiles/r2_ncu_traffic.json from this round's `ncu --set full` reports: DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) and duration of the captured launch, per kernel.
bench.py reads the JSON for roofline.traffic (scaled by the rank's share of the detections).
usage: ncu_traffic.py name=report.ncu-rep[:detections] ...      (run where ncu is installed)"""
import csv, io, json, os, subprocess, sys
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
out = {}
for arg in sys.argv[1:]:
    name, rest = arg.split('=')
    rep, _, ndet = rest.partition(':')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], zip(rows[1], rows[2])))
    rd = float(d['dram__bytes_read.sum'][1]) * UNIT[d['dram__bytes_read.sum'][0]]
    wr = float(d['dram__bytes_write.sum'][1]) * UNIT[d['dram__bytes_write.sum'][0]]
    ms = float(d['gpu__time_duration.sum'][1]) * {'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[d['gpu__time_duration.sum'][0]]
    out[name] = {'dram_bytes': rd + wr, 'dram_read': rd, 'dram_write': wr, 'ncu_ms': ms, 'detections': int(ndet or 64000001),
                 'report': os.path.basename(rep), 'kernel': rows[2][rows[0].index('Kernel Name')] if 'Kernel Name' in rows[0] else ''}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'r2_ncu_traffic.json')
json.dump(out, open(path, 'w'), indent=1)
print(json.dumps(out, indent=1))
