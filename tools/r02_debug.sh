#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
W="suzanne teapot"
echo "== default (small top)"; timeout 300 python tools/quick_bench.py $W 2>&1 | tail -2
echo "== FW_SMALL_TOP=0"; FW_SMALL_TOP=0 timeout 300 python tools/quick_bench.py $W 2>&1 | tail -2
timeout 300 python tools/r02_determinism.py teapot 1920 1080 16 | head -4
FW_SMALL_TOP=0 timeout 300 python tools/r02_determinism.py teapot 1920 1080 16 | head -2
