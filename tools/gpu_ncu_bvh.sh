#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'extend_pass' -s 2 -c 2 -f -o gpurun_out/src_teapot3 python tools/prof_run.py teapot 1920 1080 4 > gpurun_out/ncu_src_teapot3.log 2>&1; echo teapot=$?
