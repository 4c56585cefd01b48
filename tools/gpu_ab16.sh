#!/bin/bash
set -u
echo "== no prefetch"; FW_LIB_PATH=$PWD/firework_b200/libfw_nopf.so python tools/quick_bench.py random_spheres cornell_box teapot part2_all earth 2>&1 | tail -5
echo "== prefetch"; python tools/quick_bench.py random_spheres cornell_box teapot part2_all earth 2>&1 | tail -5
