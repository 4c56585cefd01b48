#!/bin/bash
# A/B: committed lib (libfw_base.so) vs working tree; then source-level ncu captures
set -u
mkdir -p gpurun_out
echo "== base"; FW_LIB_PATH=$PWD/firework_b200/libfw_base.so python tools/quick_bench.py cornell_box earth hdri_test 2>&1 | tail -4
echo "== new";  python tools/quick_bench.py cornell_box earth hdri_test 2>&1 | tail -4
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'extend_linear|shade_scatter' -c 4 -f -o gpurun_out/src_cornell python tools/prof_run.py cornell_box 300 300 64 > gpurun_out/ncu_src_cornell.log 2>&1; echo ncu_cornell=$?
$NCU -k regex:'extend_bvh_simple|shade_scatter' -s 4 -c 4 -f -o gpurun_out/src_rs python tools/prof_run.py random_spheres 960 540 8 > gpurun_out/ncu_src_rs.log 2>&1; echo ncu_rs=$?
ls -la gpurun_out/*.ncu-rep
