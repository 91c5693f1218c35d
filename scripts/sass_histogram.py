"""Opcode histogram of the shipped library's SASS per kernel (runs without a GPU):
    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt
UTC*MMA = tcgen05.mma, UTMALDG = TMA loads, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, LDGSTS = cp.async, REDUX = redux.sync (profiling recipe, "What proves a
Blackwell-native kernel")."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'iterseg_b200', 'libiterseg_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'LDGSTS', 'REDUX',
         'HMMA', 'HGMMA', 'DADD', 'DMUL', 'DFMA', 'ATOMS', 'ATOMG', 'RED', 'SHFL', 'LDS', 'STS', 'LDG', 'STG']
kern = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = subprocess.run(['cu++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = name.replace('(int)', '').replace('(bool)', '')
        kern = re.sub(r'\(.*', '', name).replace('void ', '').replace('isg::', '')
        n = 2
        base = kern
        while kern in hist:                              # distinct instantiations that print alike
            kern = f'{base} #{n}'
            n += 1
        hist[kern] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and kern:
        op = m.group(1).split('.')[0]
        hist[kern]['_total'] += 1
        if op in WATCH:
            hist[kern][op] += 1
print('arch:', re.findall(r'arch = (\S+)', out)[:1], ' kernels:', len(hist))
tot = collections.Counter()
for k, h in hist.items():
    tot.update(h)
print('TOTAL   ' + '  '.join(f'{op} {tot[op]}' for op in WATCH if tot[op]))
print()
for k, h in sorted(hist.items(), key=lambda kv: -(kv[1]['UTCHMMA'] * 1000 + kv[1]['_total'])):
    if h['_total'] < 40:
        continue
    print(f'{k[:70]:70s} {h["_total"]:6d} instr  ' + '  '.join(f'{op} {h[op]}' for op in WATCH if h[op]))
