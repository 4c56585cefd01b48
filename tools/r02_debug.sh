#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --per-config-steps 2 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo bench=$?; tail -c 300 gpurun_out/bench_quick.err
