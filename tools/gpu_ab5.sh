#!/bin/bash
set -u
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
echo "== shade6"; FW_LIB_PATH=$PWD/firework_b200/libfw_shade6.so python tools/quick_bench.py 2>&1 | tail -9
echo "== default (8)"; python tools/quick_bench.py 2>&1 | tail -9
