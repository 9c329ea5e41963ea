#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list and one --set full capture.
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -c 1500 $OUT/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
# one full capture of the hot kernel in the steady state of the rollout (exact schedule: the work of the launch is the
# same in every replay pass), one of the crossing config, one of the raster
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_substeps -c 1 \
    -o $OUT/k_substeps_full -f python tools/profile_rollout.py 4096 50 3 24 0 > $OUT/profile_rollout.log 2>&1; echo "ncu full rc=$?"
B2S_CFG=crossing timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_substeps -c 1 \
    -o $OUT/k_substeps_crossing_full -f python tools/profile_rollout.py 4096 20 3 8 0 > $OUT/profile_rollout_crossing.log 2>&1; echo "ncu crossing rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_shade -c 1 -s 3 \
    -o $OUT/k_render_full -f python tools/bench_render.py 2048 128 5 > $OUT/bench_render_under_ncu.log 2>&1; echo "ncu render rc=$?"
timeout 300 python tools/bench_render.py 2048 128 20 > $OUT/bench_render.json 2>&1
timeout 300 python tools/profile_rollout.py 4096 250 4 24 1 > $OUT/profile_rollout_free.log 2>&1
tail -n 5 $OUT/profile_rollout.log; tail -n 5 $OUT/profile_rollout_free.log
tail -3 $OUT/pytest_gpu.log
