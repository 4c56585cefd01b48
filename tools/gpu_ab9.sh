#!/bin/bash
set -u
for v in leaf1 leaf2 leaf8; do echo "== $v"; FW_LIB_PATH=$PWD/firework_b200/libfw_$v.so python tools/quick_bench.py random_spheres suzanne teapot part2_all 2>&1 | tail -4; done
echo "== default (4)"; python tools/quick_bench.py random_spheres suzanne teapot part2_all 2>&1 | tail -4
echo "== default (4), lock-step pass2"; FW_REFILL_LANES=0 python tools/quick_bench.py suzanne teapot 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
