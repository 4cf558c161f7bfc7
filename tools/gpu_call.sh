mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02z_pytest.log
( timeout 200 python tools/fuzz_parity.py 150 4242 2>&1 | tail -5 ) > gpurun_out/r02z_fuzz.log
{
for cfg in c1x64 "c4 100000"; do
  echo "== $cfg default"; RUN_REPS=4 timeout 120 python tools/run_config.py $cfg 2>&1 | tail -2
  echo "== $cfg WC=64"; SWB_TRACE_WC=64 RUN_REPS=4 timeout 120 python tools/run_config.py $cfg 2>&1 | tail -2
  echo "== $cfg NB=2"; SWB_TRACE_NB=2 RUN_REPS=4 timeout 120 python tools/run_config.py $cfg 2>&1 | tail -2
  echo "== $cfg WC=64 NB=2"; SWB_TRACE_WC=64 SWB_TRACE_NB=2 RUN_REPS=4 timeout 120 python tools/run_config.py $cfg 2>&1 | tail -2
done
echo "== c5 51M"; RUN_REPS=2 timeout 200 python tools/run_config.py c5 51000000 2>&1 | tail -2
} > gpurun_out/r02z_timings.log 2>&1
timeout 900 python bench.py > gpurun_out/r02z_bench_1gpu.json 2> gpurun_out/r02z_bench_1gpu.err
timeout 300 python bench.py --impl reference > gpurun_out/r02z_bench_reference_arm.json 2> gpurun_out/r02z_bench_reference_arm.err
tail -3 gpurun_out/r02z_pytest.log gpurun_out/r02z_fuzz.log; cat gpurun_out/r02z_timings.log; cut -c1-400 gpurun_out/r02z_bench_1gpu.json
