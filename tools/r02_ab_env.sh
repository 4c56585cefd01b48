#!/bin/bash
# Developer tool: A/B one environment switch (default library) on the quick workloads, two repetitions each.
#   VAR=FW_FUSED_SHADE A=0 B=1 bash tools/r02_ab_env.sh
mkdir -p gpurun_out
WL="${WL:-part2_all random_spheres cornell_box suzanne teapot earth hdri_test}"
{
for rep in 1 2; do
  for v in $A $B; do
    echo "== $VAR=$v (rep $rep)"; env $VAR=$v python tools/quick_bench.py $WL
  done
done
} 2>&1 | tee gpurun_out/ab_env.log
