#!/bin/bash
# A/B of variant builds on the bench workload: bash tools/variants_sweep.sh name1 name2 ...   ("default" = libb2s.so)
mkdir -p gpurun_out/variants
for v in "$@"; do
  if [ "$v" = default ]; then unset B2S_LIB; else export B2S_LIB=$PWD/robovat_b200/csrc/variants/libb2s_$v.so; fi
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e ${SWEEP_ARGS:---no-extras} > gpurun_out/variants/$v.json 2> gpurun_out/variants/$v.err
  python -c "
import json;d=json.loads(open('gpurun_out/variants/$v.json').read().strip().split('\n')[-1]);print('$v',round(d['value']/1e6,2),'M', round(d['ms_per_step'],1),'ms', 'crossing', (d.get('other_configs') or {}).get('crossing_4096',{}).get('value'))"
done
