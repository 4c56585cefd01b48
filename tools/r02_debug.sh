#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "part2_final" 2>&1 | grep -E "part2_final|passed|failed|Error|assert" | head -20
bash tools/r02_profile_all.sh
