mkdir -p gpurun_out/warps
for w in 16 14 12 15; do
  B2S_WARPS=$w timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/warps/w$w.json 2> gpurun_out/warps/w$w.err
  python -c "
import json;d=json.loads(open('gpurun_out/warps/w$w.json').read().strip().split('\n')[-1]);print('warps',$w,d['value'],d['ms_per_step'])"
done
