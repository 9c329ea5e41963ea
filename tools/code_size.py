#!/usr/bin/env python
"""Static SASS size of k_substeps by device function (CPU only). usage: python tools/code_size.py [libb2s.so]"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.join(ROOT, 'robovat_b200/csrc/libb2s.so')
with tempfile.TemporaryDirectory() as d:
    subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if 'b2s_step' in f][0]
    sass = subprocess.run(['nvdisasm', '-c', os.path.join(d, cubin)], capture_output=True, text=True).stdout
cur, cnt = 'k_substeps (body)', collections.Counter()
for l in sass.split('\n'):
    m = re.match(r'\$_Z10k_substepsILb0EEviiffi\w*\$(\w+):', l)
    if m:
        cur = m.group(1)
    if re.match(r'\s+/\*[0-9a-f]+\*/', l):
        cnt[cur] += 1
for k, v in cnt.most_common():
    print('%6d %7.1f KB  %s' % (v, v * 16 / 1024, k))
print('%6d %7.1f KB  total' % (sum(cnt.values()), sum(cnt.values()) * 16 / 1024))
