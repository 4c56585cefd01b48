#!/bin/bash
# First verification on a B200 box (gpurun --gpus 2): GPU tests incl. the multi-GPU ones, smoke, a short bench at N=1 and N=2.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo smoke=$?
timeout 900 python bench.py --steps 2 --warmup 3 --per-config-steps 2 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench1=$?; tail -c 600 gpurun_out/bench_n1.err
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --per-config-steps 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo bench2=$?; tail -c 600 gpurun_out/bench_n2.err
fi
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref=$?
