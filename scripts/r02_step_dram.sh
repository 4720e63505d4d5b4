#!/bin/bash
# whole-step DRAM traffic: dram__bytes_read/write of every launch of the bench command (ncu, cold L2 per launch = an upper bound
# on what the step moves with the 126 MB L2 warm), summed over one step by scripts/summarize_dram.py
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --blocks 1 --no-extra --no-cpu-baseline"
timeout 200 $BENCH > gpurun_out/r02_plain_bench2.log 2>&1 && \
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv \
  --log-file gpurun_out/r02_step_dram.csv $BENCH > gpurun_out/r02_ncu_dram.log 2>&1
echo "dram list rc=$? lines=$(grep -c dram__bytes_read gpurun_out/r02_step_dram.csv)"
