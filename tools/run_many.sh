#!/bin/bash
# parity + timing of the default library and every variant
for lib in robovat_b200/csrc/libb2s.so robovat_b200/csrc/variants/*.so; do [ -f $lib ] || continue
  echo "=== $lib"
  B2S_LIB=$PWD/$lib timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "push_action or stacked" 2>&1 | tail -1
  B2S_LIB=$PWD/$lib timeout -s KILL 120 python tools/profile_step.py 4096 100 4 600 2>&1 | grep "ms per"
done
