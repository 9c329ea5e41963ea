#!/bin/bash
# usage: bash tools/run_one.sh LIBNAME "pytest -k expression"
B2S_LIB=$PWD/robovat_b200/csrc/variants/libb2s_$1.so timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$2" 2>&1 | tail -5
B2S_LIB=$PWD/robovat_b200/csrc/variants/libb2s_$1.so timeout -s KILL 120 python tools/profile_step.py 4096 100 4 600 2>&1 | grep "ms per"
