#!/usr/bin/env python
"""SASS instruction count of a source line range of b2s_step.cu (CPU only). usage: python tools/loop_size.py FIRST LAST [lib]"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
first, last = int(sys.argv[1]), int(sys.argv[2])
lib = os.path.abspath(sys.argv[3]) if len(sys.argv) > 3 else os.path.join(ROOT, 'robovat_b200/csrc/libb2s.so')
with tempfile.TemporaryDirectory() as d:
    subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if 'b2s_step' in f][0]
    sass = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cubin)], capture_output=True, text=True).stdout
cur, rows = None, []
for l in sass.split('\n'):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]+\*/', l):
        rows.append(cur)
idx = [i for i, c in enumerate(rows) if c and c[0] == 'b2s_step.cu' and first <= c[1] <= last]
print('lines %d-%d: %d instructions attributed, span %d instructions (%.1f KB)' % (first, last, len(idx), max(idx) - min(idx) + 1, (max(idx) - min(idx) + 1) * 16 / 1024.0))
