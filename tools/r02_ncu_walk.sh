#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'walk_mesh' -s 1 -c 2 -f -o gpurun_out/r02_walk2_teapot python tools/prof_run.py teapot 1920 1080 8 > gpurun_out/ncu_walk_teapot.log 2>&1; echo teapot=$?
