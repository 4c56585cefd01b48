#!/bin/bash
# Quick verification of the working tree on one B200: GPU tests, smoke, default bench line, part2 bench line.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo smoke=$?
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo bench=$?
timeout 600 python bench.py --steps 2 --warmup 3 --workload part2_all --spp 128 > gpurun_out/bench_part2_all.json 2>gpurun_out/bench_part2_all.err; echo part2=$?
