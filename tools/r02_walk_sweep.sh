#!/bin/bash
# Walk-kernel verification + variant sweep on one B200: GPU tests with the walk kernels on, then the quick bench of the
# BVH workloads for the lock-step kernels (FW_WALK=0) and for each walk build in firework_b200/variants/.
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "low_spp" > gpurun_out/pytest_quick.log 2>&1; echo quick=$?; tail -15 gpurun_out/pytest_quick.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -5 gpurun_out/pytest_gpu.log
W="suzanne teapot"
echo "== lock-step (FW_WALK=0)"; FW_WALK=0 timeout 300 python tools/quick_bench.py $W 2>&1 | tail -4
echo "== walk default"; timeout 300 python tools/quick_bench.py $W 2>&1 | tail -4
for lib in firework_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"; FW_LIB_PATH=$lib timeout 300 python tools/quick_bench.py $W 2>&1 | tail -4
done
