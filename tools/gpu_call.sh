mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02ab_pytest.log
( timeout 200 python tools/fuzz_parity.py 200 31337 2>&1 | tail -5 ) > gpurun_out/r02ab_fuzz.log
timeout 900 python bench.py > gpurun_out/r02ab_bench_1gpu.json 2> gpurun_out/r02ab_bench_1gpu.err
timeout 300 python bench.py --impl reference > gpurun_out/r02ab_bench_reference_arm.json 2> gpurun_out/r02ab_bench_reference_arm.err
SWB_DEBUG_STAGE=1 timeout 300 python bench.py --steps 3 --no-configs --no-cpu-baseline > gpurun_out/r02ab_bench_stage.json 2> gpurun_out/r02ab_bench_stage.err
cat gpurun_out/r02ab_pytest.log gpurun_out/r02ab_fuzz.log; grep "free classes\|upload classes" gpurun_out/r02ab_bench_stage.err | sort -k5 -n | tail -4
python - <<'P'
import json
for f in ['gpurun_out/r02ab_bench_1gpu.json','gpurun_out/r02ab_bench_stage.json']:
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['e2e']['parts'])
P
