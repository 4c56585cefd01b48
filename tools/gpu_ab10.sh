#!/bin/bash
set -u
echo "== if-if"; FW_LIB_PATH=$PWD/firework_b200/libfw_ifif.so python tools/quick_bench.py random_spheres suzanne teapot part2_all 2>&1 | tail -4
echo "== while-while"; python tools/quick_bench.py random_spheres suzanne teapot part2_all 2>&1 | tail -4
