R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $R --nproc-per-node 8 --master-port 29558 bench.py --gpus 8 --steps 500 --warmup 10 > gpurun_out/r2_bench_n8f.json 2> gpurun_out/r2_bench_n8f.err; head -c 330 gpurun_out/r2_bench_n8f.json; echo
timeout 200 $R --nproc-per-node 8 --master-port 29568 bench.py --gpus 8 --config C5 --steps 50 --warmup 5 > gpurun_out/r2_bench_c5_n8f.json 2> gpurun_out/r2_bench_c5_n8f.err; head -c 330 gpurun_out/r2_bench_c5_n8f.json; echo
timeout 240 $R --nproc-per-node 4 --master-port 29554 bench.py --gpus 4 --steps 500 --warmup 10 > gpurun_out/r2_bench_n4f.json 2> gpurun_out/r2_bench_n4f.err; head -c 330 gpurun_out/r2_bench_n4f.json; echo
