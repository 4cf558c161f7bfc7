mkdir -p gpurun_out
{ for v in 0 1 0 1; do echo "== SWB_QS_PROF16=$v"; SWB_QS_PROF16=$v RUN_REPS=4 timeout 100 python tools/run_config.py c4 100000 2>&1 | tail -2; done; } > gpurun_out/r02ae_ab.log 2>&1
python - > gpurun_out/r02ae_choice.txt <<'P'
import re
cur=None; t={0:[],1:[]}
for l in open('gpurun_out/r02ae_ab.log'):
    m=re.match(r'== SWB_QS_PROF16=(\d)',l)
    if m: cur=int(m.group(1)); continue
    m=re.search(r'pass1 ([0-9.]+)',l)
    if m and cur is not None: t[cur].append(float(m.group(1)))
ok = t[0] and t[1] and min(t[1]) < 0.98*min(t[0])
print(1 if ok else 0)
P
export SWB_QS_PROF16=$(cat gpurun_out/r02ae_choice.txt)
echo "chosen SWB_QS_PROF16=$SWB_QS_PROF16"
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02ae_pytest.log
( timeout 200 python tools/fuzz_parity.py 200 2718 2>&1 | tail -5 ) > gpurun_out/r02ae_fuzz.log
timeout 900 python bench.py > gpurun_out/r02ae_bench_1gpu.json 2> gpurun_out/r02ae_bench_1gpu.err
cat gpurun_out/r02ae_ab.log; tail -n 3 gpurun_out/r02ae_pytest.log gpurun_out/r02ae_fuzz.log; tail -n 3 gpurun_out/r02ae_bench_1gpu.err; cut -c1-300 gpurun_out/r02ae_bench_1gpu.json
