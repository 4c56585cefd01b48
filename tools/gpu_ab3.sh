#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -5 gpurun_out/pytest_gpu.log
echo "== base"; FW_LIB_PATH=$PWD/firework_b200/libfw_base.so python tools/quick_bench.py 2>&1 | tail -9
echo "== new";  python tools/quick_bench.py 2>&1 | tail -9
