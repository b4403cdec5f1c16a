#!/usr/bin/env python
"""Per-kernel SASS opcode counts of the built library (cuobjdump -sass, no GPU needed): static instruction count, the bulk-copy engine
(UBLKCP = cp.async.bulk, SYNCS = mbarrier), global / shared loads and stores, integer dot products (IDP = __dp4a / __dp2a), POPC,
VABSDIFF4 and FP64 arithmetic.  Usage: python tools/sass_summary.py [libhvofront.so] > profiles/rNN_sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else glob.glob(os.path.join(ROOT, 'a-low*', 'libhvofront.so'))[0]
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sass.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line) if cur else None
    if m:
        cnt[cur][m.group(1).split('.')[0]] += 1
names = list(cnt)
dem = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.strip().split('\n')
rows = []
for f, d in zip(names, dem):
    c = cnt[f]
    rows.append((d.split('(')[0].replace('void ', ''), sum(c.values()), c['UBLKCP'], c['SYNCS'], c['LDG'], c['LDS'], c['STS'], c['IDP'] + c['IDP4A'], c['POPC'],
                 c['VABSDIFF4'], c['DADD'] + c['DMUL'] + c['DFMA']))
rows.sort(key=lambda r: -r[1])
print('%-36s %7s %6s %6s %5s %5s %5s %5s %5s %9s %6s' % ('kernel (sm_100a)', 'SASS', 'UBLKCP', 'SYNCS', 'LDG', 'LDS', 'STS', 'IDP', 'POPC', 'VABSDIFF4', 'FP64'))
for r in rows:
    print('%-36s %7d %6d %6d %5d %5d %5d %5d %5d %9d %6d' % r)
