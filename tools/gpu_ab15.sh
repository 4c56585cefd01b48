#!/bin/bash
set -u
for v in blk64 blk256; do echo "== $v"; FW_LIB_PATH=$PWD/firework_b200/libfw_$v.so python tools/quick_bench.py random_spheres cornell_box teapot part2_all 2>&1 | tail -4; done
echo "== default 128"; python tools/quick_bench.py random_spheres cornell_box teapot part2_all 2>&1 | tail -4
