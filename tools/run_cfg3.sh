#!/bin/bash
timeout -s KILL 400 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== cfg2 mid-push"
timeout -s KILL 120 python tools/profile_step.py 4096 100 4 600 2>&1 | grep "ms per"
echo "=== cfg3 crossing, 8 concave movables"
B2S_CFG=crossing timeout -s KILL 300 python tools/profile_step.py 4096 50 3 600 2>&1 | tail -4
echo "=== bench"
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value %.3fM ms %.1f launch ms %.2f' % (d['value']/1e6,d['ms_per_step'],d['roofline']['avg_launch_ms']))"
