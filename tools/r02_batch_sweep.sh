#!/bin/bash
# Developer tool: batch-size sweep at bench sizes (device-resident value only).
mkdir -p gpurun_out
for wl in teapot part2_all; do
  for b in 67108864 134217728 201326592 268435456; do
    echo "== $wl FW_BATCH_PATHS=$b"
    FW_BATCH_PATHS=$b python bench.py --workload $wl --steps 3 --warmup 1 --no-per-config --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'launches', d['gpu_launches'])"
  done
done 2>&1 | tee gpurun_out/batch_sweep.log
