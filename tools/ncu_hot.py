#!/usr/bin/env python
"""Hot spots of one kernel from an .ncu-rep (SASS view): python tools/ncu_hot.py rep [kernel-regex] [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else '.'; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:200])
rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
data = []
for r in rows[1:]:
    if len(r) != len(h) or r[0] == 'Address': break
    data.append(r)
tot = sum(int(r[ix['# Samples']]) for r in data)
texec = sum(int(r[ix['Instructions Executed']]) for r in data)
print('total samples', tot, 'warp instr executed', texec, 'SASS lines', len(data))
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print('stall totals:', {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.01 * tot})
top = sorted(enumerate(data), key=lambda x: -int(x[1][ix['# Samples']]))[:N]
for i, r in sorted(top):
    st = {s[6:]: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) > 0.15 * int(r[ix['# Samples']])}
    print('%5d %6.2f%% exec %9s  %-70s %s' % (i, 100. * int(r[ix['# Samples']]) / tot, r[ix['Instructions Executed']], r[ix['Source']].strip()[:70], st))
