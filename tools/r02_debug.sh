#!/bin/bash
set -u
W="cornell_box random_spheres part2_all suzanne teapot"
echo "== default"; timeout 300 python tools/quick_bench.py $W 2>&1 | tail -5
for lib in firework_b200/variants/*.so; do echo "== $lib"; FW_LIB_PATH=$lib timeout 300 python tools/quick_bench.py $W 2>&1 | tail -5; done
