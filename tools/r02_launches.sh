#!/bin/bash
# ncu launch lists: (1) the bench command itself (first 400 launches, gpu__time_duration only — the contract's pass);
# (2) one render of each heavy config with busy lanes and IPC per launch, summarised by tools/launch_shares.py.
set -u
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-per-config --no-cpu-baseline > gpurun_out/launches_bench_plain.json 2> gpurun_out/launches_bench_plain.err || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-per-config --no-cpu-baseline > gpurun_out/launches_bench_ncu.log 2>&1; echo "bench list rc=$?"
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_active
run() {
  ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_$1.csv python tools/prof_run.py $1 $2 $3 $4 > gpurun_out/launches_$1.log 2>&1; echo "$1 rc=$?"
}
run part2_all 3840 2160 4
run cornell_box 300 300 256
run random_spheres 960 540 32
run suzanne 1920 1080 16
run teapot 1920 1080 16
python tools/launch_shares.py gpurun_out/launches_part2_all.csv gpurun_out/launches_cornell_box.csv gpurun_out/launches_random_spheres.csv gpurun_out/launches_suzanne.csv gpurun_out/launches_teapot.csv > gpurun_out/r02_launch_shares.txt 2>&1
head -c 400 gpurun_out/r02_launch_shares.txt
