#!/bin/bash
# A/B of the optional cluster-multicast path of gconv_tc_kernel (MMVAE_MC_MIN_CHUNKS, off by default): parity self-test,
# per-layer timeline of the deep-K layers and the step time with the option on (6) and off (0).
TAG=$1
L="encoder.layer4.0.conv2 encoder.layer3.0.conv2 encoder.layer4.0.conv1 decoder.uplayer1.0.conv2 encoder.layer2.0.conv2"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tc_selftest.py -x -q > gpurun_out/${TAG}_selftest.log 2>&1; echo "selftest rc=$?"; tail -n 3 gpurun_out/${TAG}_selftest.log
for mc in 6 0; do
  echo "---- MC_MIN_CHUNKS=$mc"
  MMVAE_MC_MIN_CHUNKS=$mc timeout 100 python scripts/trace_conv.py $L > gpurun_out/${TAG}_trace_$mc.log 2>&1; echo "rc=$?"
  grep "==" gpurun_out/${TAG}_trace_$mc.log | cut -c1-120
  MMVAE_MC_MIN_CHUNKS=$mc timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | cut -c1-260
done
