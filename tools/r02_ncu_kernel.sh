#!/bin/bash
# Developer tool: one `ncu --set full` capture of one kernel + per-line table.
#   bash tools/r02_ncu_kernel.sh <scene> <kernel regex> <skip> <tag> [w h spp]
scene=$1; kern=$2; skip=$3; tag=$4; shift 4
mkdir -p gpurun_out
python tools/prof_run.py $scene "$@" > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/$tag \
    python tools/prof_run.py $scene "$@" > gpurun_out/${tag}_ncu.log 2>&1
python tools/ncu_lines.py gpurun_out/$tag.ncu-rep 1 80 > gpurun_out/${tag}_lines.txt 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
