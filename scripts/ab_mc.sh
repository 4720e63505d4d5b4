#!/bin/bash
# A/B of the deep-K options of gconv_tc_kernel (cluster multicast of the A boxes, rotating TMEM accumulators)
TAG=$1
L="encoder.layer4.0.conv2 encoder.layer3.0.conv2 encoder.layer4.0.conv1 decoder.uplayer1.0.conv2 encoder.layer2.0.conv2"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tc_selftest.py -x -q > gpurun_out/${TAG}_selftest.log 2>&1; echo "selftest rc=$?"; tail -n 3 gpurun_out/${TAG}_selftest.log
for cfg in "6 3" "0 3" "6 0" "0 0"; do
  set -- $cfg
  echo "---- MC_MIN_CHUNKS=$1 NACC_MIN_CHUNKS=$2"
  MMVAE_MC_MIN_CHUNKS=$1 MMVAE_NACC_MIN_CHUNKS=$2 timeout 100 python scripts/trace_conv.py $L > gpurun_out/${TAG}_trace_$1_$2.log 2>&1; echo "rc=$?"
  grep "==" gpurun_out/${TAG}_trace_$1_$2.log | cut -c1-120
  MMVAE_MC_MIN_CHUNKS=$1 MMVAE_NACC_MIN_CHUNKS=$2 timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | cut -c1-260
done
