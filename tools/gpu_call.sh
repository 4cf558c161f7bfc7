mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 110 $TR --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02af_bench_2gpu.json 2> gpurun_out/r02af_bench_2gpu.err; echo "default rc=$?"
timeout 80 $TR --master-port 29512 bench.py --gpus 2 --config c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02af_bench_c5_2gpu.json 2> gpurun_out/r02af_bench_c5_2gpu.err; echo "c5 rc=$?"
timeout 60 $TR --master-port 29513 bench.py --gpus 2 --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02af_bench_c4_2gpu.json 2> gpurun_out/r02af_bench_c4_2gpu.err; echo "c4 rc=$?"
for f in gpurun_out/r02af_bench_*json; do echo $f; cut -c1-200 $f; done; tail -n 2 gpurun_out/r02af_*.err
