set -u
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
for s in suzanne teapot; do
  ncu --metrics $M --clock-control none -k regex:'extend|walk_mesh|classify_mesh' -c 33 -f -o gpurun_out/r02_extend_$s python tools/prof_run.py $s 1920 1080 16 > gpurun_out/ncu_extend_$s.log 2>&1; echo "$s rc=$?"
  python tools/ncu_traffic.py gpurun_out/r02_extend_$s.ncu-rep gpurun_out/r02_extend_traffic_$s.json > /dev/null 2>gpurun_out/traffic_$s.err
  rm -f gpurun_out/r02_extend_$s.ncu-rep
done
bash tools/r02_launches.sh > gpurun_out/launches.log 2>&1
tail -3 gpurun_out/launches.log
