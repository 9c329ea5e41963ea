#!/bin/bash
B2S_CFG=crossing B2S_LIB=$PWD/robovat_b200/csrc/variants/libb2s_prof.so timeout -s KILL 300 python tools/profile_step.py 4096 50 3 600 2>&1 | tail -12
