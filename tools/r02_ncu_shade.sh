#!/bin/bash
# Developer tool: one `ncu --set full` capture of the Lambertian shade kernel (cornell_box, bounce 1) + per-line table.
mkdir -p gpurun_out
python tools/prof_run.py cornell_box 300 300 256 > gpurun_out/shade_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:shade_scatter_kernel -s 1 -c 1 -f -o gpurun_out/r02_shade_cornell \
    python tools/prof_run.py cornell_box 300 300 256 > gpurun_out/shade_ncu.log 2>&1
python tools/ncu_lines.py gpurun_out/r02_shade_cornell.ncu-rep 1 70 > gpurun_out/r02_shade_cornell_lines.txt 2>&1
ncu -i gpurun_out/r02_shade_cornell.ncu-rep --page raw --csv > gpurun_out/r02_shade_cornell_raw.csv 2>/dev/null
ls -la gpurun_out/
