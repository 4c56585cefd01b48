#!/bin/bash
# GPU tests + the quick device-timed bench of every workload (developer loop)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -2 gpurun_out/pytest_gpu.log
python tools/quick_bench.py 2>&1 | tail -9
