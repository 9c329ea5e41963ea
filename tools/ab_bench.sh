#!/bin/bash
# A/B of the default library and every variant on one box: mid-push launch time and the bench line (no CPU leg)
# usage: gpurun --timeout 900 -- 'bash tools/ab_bench.sh OUTFILE'
OUT=${1:-gpurun_out/ab.txt}
for lib in robovat_b200/csrc/libb2s.so robovat_b200/csrc/variants/*.so robovat_b200/csrc/libb2s.so; do [ -f $lib ] || continue
  echo "== $lib" >> $OUT
  B2S_LIB=$PWD/$lib timeout -s KILL 120 python tools/profile_step.py 4096 100 4 600 2>&1 | grep "ms per" >> $OUT
  B2S_LIB=$PWD/$lib timeout -s KILL 250 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])" >> $OUT
done
cat $OUT
