#!/bin/bash
# copy the artefacts of a tools/gpu_check.sh run (gpurun_out/TAG) into profiles/ under the round prefix
TAG=${1:-r02}; PRE=${2:-r02}
SRC=gpurun_out/$TAG
python tools/ncu_summary.py $SRC/k_substeps_full.ncu-rep profiles/${PRE}_k_substeps_ncu_full_summary.txt > /dev/null
cp $SRC/bench.json profiles/${PRE}_bench_n1.json
cp $SRC/bench_ref.json profiles/${PRE}_bench_reference_n1.json
cp $SRC/launches.csv profiles/${PRE}_launches.csv
cp $SRC/profile_rollout.log profiles/${PRE}_profile_rollout.log
cp $SRC/profile_rollout_free.log profiles/${PRE}_profile_rollout_free.log
cp $SRC/profile_rollout_crossing.log profiles/${PRE}_profile_rollout_crossing.log
cp $SRC/bench_render.json profiles/${PRE}_bench_render.json
python tools/ncu_summary.py $SRC/k_substeps_crossing_full.ncu-rep profiles/${PRE}_k_substeps_crossing_ncu_full_summary.txt > /dev/null
python tools/ncu_summary.py $SRC/k_render_full.ncu-rep profiles/${PRE}_k_render_shade_ncu_full_summary.txt > /dev/null
cp $SRC/pytest_gpu.log profiles/${PRE}_pytest_gpu.log
cp $SRC/smoke.log profiles/${PRE}_smoke.log
python - <<PY
import csv, io, subprocess, json
rep='$SRC/k_substeps_full.ncu-rep'
raw=list(csv.reader(io.StringIO(subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout)))
h,u,v=raw[0],raw[1],raw[2]
def g(k):
    x=float(v[h.index(k)]); un=u[h.index(k)]
    return x*{'Mbyte':1e6,'Kbyte':1e3,'Gbyte':1e9,'byte':1}.get(un,1)
rd,wr=g('dram__bytes_read.sum'),g('dram__bytes_write.sum')
envsub=4096*50
out={'kernel':'k_substeps','source':'ncu --set full --clock-control none --import-source on, tools/profile_rollout.py 4096 50 3 24 0 (4096 envs x 50 substeps in the steady state of the rollout, exact schedule); summary in profiles/${PRE}_k_substeps_ncu_full_summary.txt',
     'dram_bytes_read':rd,'dram_bytes_write':wr,'env_substeps':envsub,'dram_bytes_per_env_substep':(rd+wr)/envsub,
     'gpu_time_ms':float(v[h.index('gpu__time_duration.sum')]),'warp_instructions':float(v[h.index('smsp__inst_executed.sum')])}
json.dump(out,open('profiles/${PRE}_k_substeps_dram_traffic.json','w'),indent=1)
print(out)
PY
