#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py suzanne teapot 2>&1 | tail -2
for w in suzanne teapot; do
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_active --clock-control none --csv --log-file gpurun_out/launches_$w.csv python tools/prof_run.py $w 1920 1080 16 > /dev/null 2>&1; echo ncu_$w=$?
done
