#!/usr/bin/env python
"""Static count of local-memory instructions (LDL/STL) of k_substeps by source line (needs nvdisasm; CPU only).
usage: python tools/local_ops.py [libb2s.so]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else 'robovat_b200/csrc/libb2s.so')
with tempfile.TemporaryDirectory() as d:
    subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if 'b2s_step' in f][0]
    sass = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cubin)], capture_output=True, text=True).stdout
cur, cnt, n = None, collections.Counter(), 0
for l in sass.split('\n'):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]+\*/', l):
        n += 1
        if re.search(r'\b(LDL|STL)\b', l):
            cnt[cur] += 1
print('SASS instructions', n, ' LDL/STL', sum(cnt.values()))
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:int(os.environ.get('TOP', 25))]:
    print('%-22s %5d  %s' % ('%s:%d' % k, v, ''))
