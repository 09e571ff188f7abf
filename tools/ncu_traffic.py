#!/usr/bin/env python
"""profiles/traffic.json from ncu --set full captures: python tools/ncu_traffic.py workload rep [rep ...]
DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches) per kernel."""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
workload = sys.argv[1]
path = os.path.join(ROOT, 'profiles', 'traffic.json')
tab = json.load(open(path)) if os.path.exists(path) else {}
for rep in sys.argv[2:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out))); h = rows[0]; units = rows[1]
    acc = {}
    for r in rows[2:]:
        name = re.sub(r'<.*', '', r[h.index('Kernel Name')]).replace('void ', '').strip()
        full = r[h.index('Kernel Name')]
        b = 0.
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = h.index(m); b += float(r[i].replace(',', '')) * UNIT[units[i]]
        t = float(r[h.index('gpu__time_duration.sum')].replace(',', '')); tu = units[h.index('gpu__time_duration.sum')]
        t *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1., 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1., 's': 1e3, 'second': 1e3}[tu]
        acc.setdefault(name, []).append((b, t, full))
    for name, v in acc.items():
        tab.setdefault(workload, {})[name] = {'bytes_per_launch': sum(x[0] for x in v) / len(v), 'launches_captured': len(v),
                                              'ncu_ms_per_launch': sum(x[1] for x in v) / len(v), 'instances': sorted(set(x[2] for x in v)),
                                              'report': os.path.basename(rep)}
json.dump(tab, open(path, 'w'), indent=1, sort_keys=True)
print(json.dumps(tab.get(workload), indent=1)[:1500])
