#!/bin/bash
# one gpurun call: tcgen05-vs-SIMT self-test; bounded by `timeout` so a hung kernel cannot eat the box
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_selftest.py -x -q -s -m gpu > gpurun_out/selftest.log 2>&1
echo "exit $?" >> gpurun_out/selftest.log
tail -n 150 gpurun_out/selftest.log
