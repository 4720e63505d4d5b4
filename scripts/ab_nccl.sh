#!/bin/bash
# A/B of NCCL settings on the data-parallel headline workload (N ranks): ms per step
N=$1; shift
mkdir -p gpurun_out
port=29600
for setting in "$@"; do
  port=$((port+1))
  out=$(env $setting MMVAE_BENCH_STALL_S=120 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $N --steps 100 --warmup 10 --blocks 3 --no-extra 2>/dev/null | \
        python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  e2e %.4f  adam %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['with_adam']['ms_per_step']))")
  echo "[N=$N $setting] $out" | tee -a gpurun_out/nccl_ab.log
done
