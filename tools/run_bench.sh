#!/bin/bash
run() { timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value %.3fM ms %.1f launch ms %.2f' % (d['value']/1e6,d['ms_per_step'],d['roofline']['avg_launch_ms']))"; }
echo "default"; run
echo "EPB=28 (no spare slots, plain round-robin deal)"; B2S_EPB=28 run
echo "simplex noinline"; B2S_LIB=$PWD/robovat_b200/csrc/variants/libb2s_sxno.so run
echo "simplex noinline + EPB=28"; B2S_EPB=28 B2S_LIB=$PWD/robovat_b200/csrc/variants/libb2s_sxno.so run
echo "default again"; run
