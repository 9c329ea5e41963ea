#!/bin/bash
# quick GPU iteration: parity tests + mid-push timing (+ optional ncu full capture when $2 = ncu)
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.log
timeout -s KILL 120 python tools/profile_step.py 4096 100 4 600 2>&1 | tail -4 | tee $OUT/profile_step.log
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_substeps -c 1 \
    -o $OUT/k_substeps_full -f python tools/profile_step.py 4096 50 3 600 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
fi
