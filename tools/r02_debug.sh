#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== checks build"; FW_LIB_PATH=firework_b200/variants/walk_checks.so FW_DEBUG_SYNC=1 timeout 200 python tools/prof_run.py teapot 1920 1080 32 2>&1 | tail -2 | cut -c1-250
timeout 300 python tools/r02_determinism.py suzanne 1920 1080 32 | head -4
timeout 300 python tools/r02_determinism.py teapot 1920 1080 32 | head -4
W="suzanne teapot"
echo "== walk default"; timeout 300 python tools/quick_bench.py $W 2>&1 | tail -2
for lib in firework_b200/variants/walk_idle4.so firework_b200/variants/walk_idle16.so firework_b200/variants/walk_mb8.so; do
  echo "== $lib"; FW_LIB_PATH=$lib timeout 300 python tools/quick_bench.py $W 2>&1 | tail -2
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
