#!/bin/bash
for lib in robovat_b200/csrc/libb2s.so robovat_b200/csrc/variants/libb2s_rpt.so; do [ -f $lib ] || continue
  echo "=== $lib"; B2S_LIB=$PWD/$lib timeout -s KILL 200 python tools/bench_render.py 2048 128 20 2>&1 | tail -1
done
B2S_LIB=$PWD/robovat_b200/csrc/libb2s.so timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "render" 2>&1 | tail -1
echo "=== cfg3 crossing, 8 concave movables"
B2S_CFG=crossing timeout -s KILL 300 python tools/profile_step.py 4096 50 3 600 2>&1 | tail -4
