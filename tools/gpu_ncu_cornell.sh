#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'shade_scatter|raygen|emissive' -s 0 -c 4 -f -o gpurun_out/src_cornell5 python tools/prof_run.py cornell_box 300 300 256 > gpurun_out/ncu_src_cornell5.log 2>&1; echo ncu_cornell=$?
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cornell5.csv python tools/prof_run.py cornell_box 300 300 256 > /dev/null 2>&1; echo launches=$?
