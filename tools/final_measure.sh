#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): tests, default bench, ncu launch list + full capture of
# the extend launches of the same command, then every workload.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo smoke=$?
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_default.csv $B > gpurun_out/ncu_l.log 2>&1; echo launchlist=$?
ncu --set full --clock-control none --import-source on -k regex:extend -s 99 -c 11 -o gpurun_out/prof_final_extend_cornell $B > gpurun_out/ncu_f.log 2>&1; echo full=$?
for w in cornell_box random_spheres suzanne teapot hdri_test earth; do timeout 400 python bench.py --steps 3 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2>gpurun_out/bench_$w.err; echo $w $?; done
timeout 600 python bench.py --steps 2 --warmup 3 --workload part2_all --spp 128 > gpurun_out/bench_part2_all.json 2>gpurun_out/bench_part2_all.err; echo part2 $?
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo ref $?
