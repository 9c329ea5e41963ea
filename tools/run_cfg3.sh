#!/bin/bash
timeout -s KILL 400 python -m pytest tests -m gpu -q 2>&1 | tail -2
echo "=== cfg3 crossing, 8 concave movables (generic solve path)"
B2S_CFG=crossing timeout -s KILL 300 python tools/profile_step.py 4096 50 3 600 2>&1 | tail -4
echo "=== cfg3 with the V-HACD URDF movables"
B2S_CFG=crossing B2S_MOVABLE=vhacd timeout -s KILL 300 python tools/profile_step.py 4096 50 3 600 2>&1 | tail -4
