#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== base"; FW_LIB_PATH=$PWD/firework_b200/libfw_base.so python tools/quick_bench.py cornell_box earth hdri_test volume conics_cli 2>&1 | tail -5
echo "== new";  python tools/quick_bench.py cornell_box earth hdri_test volume conics_cli 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -5 gpurun_out/pytest_gpu.log
