#!/bin/bash
set -u
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
python tools/quick_bench.py random_spheres cornell_box part2_all volume 2>&1 | tail -4
