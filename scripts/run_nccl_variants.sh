R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
i=0
for envs in "NCCL_PROTO=Simple" "NCCL_ALGO=NVLS" "NCCL_PROTO=LL128" "NCCL_NVLS_ENABLE=1 NCCL_ALGO=NVLSTree,NVLS,Ring"; do
  i=$((i+1))
  echo "== $envs"
  env $envs timeout 120 $R --nproc-per-node 8 --master-port 2959$i bench.py --gpus 8 --config C5 --steps 40 --warmup 5 > gpurun_out/r2_nccl_$i.json 2> gpurun_out/r2_nccl_$i.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2_nccl_$i.json')); print(d['ms_per_step'], d['value'], d['parity_check']['ok'])
except Exception as e: print('failed', e)"
  tail -2 gpurun_out/r2_nccl_$i.err | cut -c1-200
done
