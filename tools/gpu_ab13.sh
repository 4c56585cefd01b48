#!/bin/bash
set -u
echo "== reference interior"; FW_BVH_SAH=0 python tools/quick_bench.py random_spheres suzanne teapot part2_all conics_cli 2>&1 | tail -5
echo "== SAH interior"; python tools/quick_bench.py random_spheres suzanne teapot part2_all conics_cli 2>&1 | tail -5
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
