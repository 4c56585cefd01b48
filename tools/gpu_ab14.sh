#!/bin/bash
set -u
echo "== min blocks 6"; FW_LIB_PATH=$PWD/firework_b200/libfw_ext6.so python tools/quick_bench.py random_spheres suzanne teapot part2_all 2>&1 | tail -4
echo "== default (8)"; python tools/quick_bench.py random_spheres suzanne teapot part2_all 2>&1 | tail -4
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
