#!/bin/bash
# Developer tool: A/B the default library against every firework_b200/variants/*.so on a few workloads (device-timed, best of 3).
mkdir -p gpurun_out
WL="${WL:-part2_all random_spheres cornell_box}"
{
for rep in 1 2; do
  echo "== default (rep $rep)"; python tools/quick_bench.py $WL
  for v in firework_b200/variants/*.so; do
    echo "== $v (rep $rep)"; FW_LIB_PATH=$v python tools/quick_bench.py $WL
  done
done
} 2>&1 | tee gpurun_out/ab.log
