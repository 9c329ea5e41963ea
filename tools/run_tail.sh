#!/bin/bash
echo "=== tail: after 3250 substeps"
B2S_LIB=$PWD/robovat_b200/csrc/variants/libb2s_prof.so timeout -s KILL 300 python tools/profile_step.py 4096 250 3 3250 2>&1 | tail -10
echo "=== bench"
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value %.3fM ms %.1f launch ms %.2f' % (d['value']/1e6,d['ms_per_step'],d['roofline']['avg_launch_ms']))"
