#!/bin/bash
set -u
for r in 0 8 16 24 28; do echo "== FW_REFILL_LANES=$r"; FW_REFILL_LANES=$r python tools/quick_bench.py suzanne teapot 2>&1 | tail -2; done
FW_REFILL_LANES=16 timeout 600 python -m pytest tests -m gpu -x -q -k "first_hit or low_spp or nan or full_size or meshes" > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
