#!/bin/bash
set -u
for s in teapot suzanne random_spheres part2_all; do echo "== $s"; FW_DEBUG_STEPS=1 python tools/prof_run.py $s 1920 1080 2 2>&1 | grep "fw debug" | cut -c1-170; done
