#!/bin/bash
timeout -s KILL 120 python tools/profile_step.py 4096 100 4 600 2>&1 | grep "ms per"
B2S_LIB=$PWD/robovat_b200/csrc/variants/libb2s_prof.so timeout -s KILL 200 python tools/profile_step.py 4096 100 4 600 2>&1 | tail -9
