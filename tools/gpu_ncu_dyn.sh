#!/bin/bash
set -u
NCU="ncu --set full --clock-control none --import-source on"
FW_REFILL_LANES=24 $NCU -k regex:'extend_pass2' -s 1 -c 1 -f -o gpurun_out/src_teapot_dyn python tools/prof_run.py teapot 1920 1080 4 > gpurun_out/ncu_src_teapot_dyn.log 2>&1; echo dyn=$?
