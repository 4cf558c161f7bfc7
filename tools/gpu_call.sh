mkdir -p gpurun_out
export RUN_REPS=1
T0=$(date +%s)
M=sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active
SEC="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section LaunchStats --section Occupancy"
full() { # name, kernel regex, count, which, args...
  name=$1; k=$2; c=$3; which=$4; shift 4
  timeout 240 ncu --set full --clock-control none -f -k regex:$k -c $c -o gpurun_out/r02ad_$name python tools/run_config.py "$@" > gpurun_out/r02ad_ncu_$name.log 2>&1
  echo "$name rc=$? t=$(( $(date +%s) - T0 ))"
  python tools/ncu_summary.py gpurun_out/r02ad_$name.ncu-rep $which > gpurun_out/r02ad_${name}_summary.txt 2>&1
  rm -f gpurun_out/r02ad_$name.ncu-rep
}
full trace_c1x64 '^trace_kernel' 1 trace_kernel c1x64
full qs_c4 'qs_' 2 qs_ c4 100000
full units_c5_2M 'score_units_kernel' 1 score_units c5 2000000
timeout 300 ncu --replay-mode application --clock-control none -f -k regex:score_units_kernel -c 1 $SEC --metrics $M \
  -o gpurun_out/r02ad_units_c5_51M python tools/run_config.py c5 51000000 > gpurun_out/r02ad_ncu_units_c5_51M.log 2>&1
echo "units 51M rc=$? t=$(( $(date +%s) - T0 ))"
python tools/ncu_summary.py gpurun_out/r02ad_units_c5_51M.ncu-rep score_units > gpurun_out/r02ad_units_c5_51M_summary.txt 2>&1; rm -f gpurun_out/r02ad_units_c5_51M.ncu-rep
timeout 240 ncu --clock-control none -f -k 'regex:^score_kernel' -c 1 $SEC --metrics $M,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
  -o gpurun_out/r02ad_score_c3 python tools/run_config.py c3 75776 > gpurun_out/r02ad_ncu_score_c3.log 2>&1
echo "score c3 rc=$? t=$(( $(date +%s) - T0 ))"
python tools/ncu_summary.py gpurun_out/r02ad_score_c3.ncu-rep score_kernel > gpurun_out/r02ad_score_c3_summary.txt 2>&1; rm -f gpurun_out/r02ad_score_c3.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ad_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02ad_ncu_launches.log 2>&1
echo "launch list rc=$? t=$(( $(date +%s) - T0 ))"
rm -f gpurun_out/*.ncu-rep; du -sh gpurun_out; head -12 gpurun_out/r02ad_units_c5_51M_summary.txt
