#!/bin/bash
set -u
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -2 gpurun_out/pytest_gpu.log
FW_BENCH_DEBUG=1 python bench.py --steps 2 --warmup 3 --workload part2_all --spp 128 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step"
for w in random_spheres earth hdri_test teapot; do python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(j['value']), round(j['e2e']['value']))"; done
