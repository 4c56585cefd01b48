#!/bin/bash
set -u
for v in shade6 shade8; do echo "== $v"; FW_LIB_PATH=$PWD/firework_b200/libfw_$v.so python tools/quick_bench.py random_spheres cornell_box teapot part2_all earth 2>&1 | tail -5; done
echo "== default (10)"; python tools/quick_bench.py 2>&1 | tail -9
echo "== no miss skip"; FW_SKIP_ZERO_MISS=0 python tools/quick_bench.py cornell_box 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
