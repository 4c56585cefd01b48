#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'extend_bvh_simple' -s 1 -c 2 -f -o gpurun_out/src_rs2 python tools/prof_run.py random_spheres 960 540 16 > gpurun_out/ncu_src_rs2.log 2>&1; echo rs=$?
$NCU -k regex:'extend_pass' -s 2 -c 2 -f -o gpurun_out/src_teapot2 python tools/prof_run.py teapot 1920 1080 4 > gpurun_out/ncu_src_teapot2.log 2>&1; echo teapot=$?
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rs2.csv python tools/prof_run.py random_spheres 960 540 32 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_teapot2.csv python tools/prof_run.py teapot 1920 1080 8 > /dev/null 2>&1
