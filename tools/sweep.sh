#!/bin/bash
# Time every tuning variant of libb2s.so (robovat_b200/csrc/variants/*.so) on the mid-push workload and check parity.
# usage: gpurun --timeout 1200 -- 'bash tools/sweep.sh TAG [envs-per-block list]'
TAG=${1:-sweep}; shift
EPBS=${@:-0}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for lib in robovat_b200/csrc/libb2s.so robovat_b200/csrc/variants/*.so; do
  [ -f $lib ] || continue
  for epb in $EPBS; do
    echo "=== $lib EPB=$epb" | tee -a $OUT/sweep.log
    B2S_LIB=$PWD/$lib B2S_EPB=$epb timeout -s KILL 120 python tools/profile_step.py 4096 100 4 600 2>&1 | tail -4 | tee -a $OUT/sweep.log
  done
  B2S_LIB=$PWD/$lib timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "push_action or crossing" 2>&1 | tail -2 | tee -a $OUT/sweep.log
done
