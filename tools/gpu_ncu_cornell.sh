#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'extend_linear|shade_scatter|miss_kernel' -s 3 -c 6 -f -o gpurun_out/src_cornell2 python tools/prof_run.py cornell_box 300 300 256 > gpurun_out/ncu_src_cornell2.log 2>&1; echo ncu_cornell=$?
