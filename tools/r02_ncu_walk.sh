#!/bin/bash
# ncu --set full (+source) of the walk kernels: random_spheres top-level walk at bounce 1-2, teapot mesh walk at bounce 1.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'walk_top' -s 1 -c 2 -f -o gpurun_out/r02_walk_rs python tools/prof_run.py random_spheres 960 540 32 > gpurun_out/ncu_walk_rs.log 2>&1; echo rs=$?
$NCU -k regex:'walk_mesh|walk_top' -s 2 -c 2 -f -o gpurun_out/r02_walk_teapot python tools/prof_run.py teapot 1920 1080 8 > gpurun_out/ncu_walk_teapot.log 2>&1; echo teapot=$?
FW_WALK=0 $NCU -k regex:'extend_bvh' -s 1 -c 2 -f -o gpurun_out/r02_lock_rs python tools/prof_run.py random_spheres 960 540 32 > gpurun_out/ncu_lock_rs.log 2>&1; echo lock_rs=$?
