#!/bin/bash
set -u
for n in 2368 3552 4736 7104 9472; do echo "== FW_SEGMENTS=$n"; FW_SEGMENTS=$n python tools/quick_bench.py random_spheres cornell_box teapot part2_all earth 2>&1 | tail -5; done
