R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 250 > gpurun_out/r2_pytest_dist8.log 2>&1; tail -1 gpurun_out/r2_pytest_dist8.log
for n in 8 4 2; do
  timeout 240 $R --nproc-per-node $n --master-port 2955$n bench.py --gpus $n --steps 500 --warmup 10 > gpurun_out/r2_bench_n${n}c.json 2> gpurun_out/r2_bench_n${n}c.err; head -c 330 gpurun_out/r2_bench_n${n}c.json; echo
  timeout 200 $R --nproc-per-node $n --master-port 2956$n bench.py --gpus $n --config C5 --steps 50 --warmup 5 > gpurun_out/r2_bench_c5_n${n}c.json 2> gpurun_out/r2_bench_c5_n${n}c.err; head -c 330 gpurun_out/r2_bench_c5_n${n}c.json; echo
done
timeout 150 $R --nproc-per-node 8 --master-port 29571 scripts/dist_timeline.py 2>&1 | grep -v "arn\|\*\*\*\|OMP" > gpurun_out/r2_dist_timeline_n8.txt
timeout 150 $R --nproc-per-node 8 --master-port 29572 scripts/bank_timeline.py 2>&1 | grep -v "arn\|\*\*\*\|OMP" > gpurun_out/r2_bank_timeline_n8.txt; head -24 gpurun_out/r2_bank_timeline_n8.txt | cut -c1-140
