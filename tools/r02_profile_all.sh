#!/bin/bash
# ncu capture of the extend kernels of ONE batch of every BASELINE.json config (the configs' own resolutions; spp = what
# one 32 Mi-path batch holds) for profiles/r02_extend_traffic_<config>.json (tools/ncu_traffic.py): DRAM bytes, busy lanes
# per instruction, IPC, pipe utilisation, cache hit rates.  A named metric list keeps this to a few replay passes per
# launch (a --set full capture of all seven configs costs 18 GPU-minutes); the headline config additionally gets a
# --set full capture of its first three extend launches.
set -u
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
run() {  # name width height spp launches
  ncu --metrics $M --clock-control none -k regex:'extend|walk_mesh|classify_mesh' -c $5 -f -o gpurun_out/r02_extend_$1 python tools/prof_run.py $1 $2 $3 $4 > gpurun_out/ncu_extend_$1.log 2>&1; echo "$1 rc=$?"
  python tools/ncu_traffic.py gpurun_out/r02_extend_$1.ncu-rep gpurun_out/r02_extend_traffic_$1.json > /dev/null 2>gpurun_out/traffic_$1.err; echo "  summary rc=$?"
  rm -f gpurun_out/r02_extend_$1.ncu-rep    # gpurun brings back at most 64 MiB
}
run part2_all 3840 2160 4 11
run random_spheres 960 540 32 11
run cornell_box 300 300 372 11
run suzanne 1920 1080 16 33
run teapot 1920 1080 16 33
run hdri_test 500 250 268 11
run earth 800 800 52 11
ncu --set full --clock-control none --import-source on -k regex:'extend' -c 3 -f -o gpurun_out/r02_full_part2_all python tools/prof_run.py part2_all 3840 2160 4 > gpurun_out/ncu_full_part2_all.log 2>&1; echo "full part2_all rc=$?"
ls -la gpurun_out/*.ncu-rep
