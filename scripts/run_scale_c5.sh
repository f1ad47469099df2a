R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 250 -k bank > gpurun_out/r2_pytest_dist8.log 2>&1; tail -1 gpurun_out/r2_pytest_dist8.log
for n in 8 4 2; do
  timeout 200 $R --nproc-per-node $n --master-port 2956$n bench.py --gpus $n --config C5 --steps 50 --warmup 5 > gpurun_out/r2_bench_c5_n${n}d.json 2> gpurun_out/r2_bench_c5_n${n}d.err; head -c 330 gpurun_out/r2_bench_c5_n${n}d.json; echo
done
