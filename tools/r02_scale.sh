#!/bin/bash
# On an 8-GPU box: multi-GPU tests, fw_render_multi bench, bench.py strong scaling at N = 2, 4, 8 (N = 1 is measured on its own box).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo pytest_multi=$?; tail -3 gpurun_out/pytest_multi.log
timeout 600 python tools/multi_bench.py part2_all 128 2>&1 | tail -8
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --per-config-steps 2 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "bench N=$n rc=$?"
done
