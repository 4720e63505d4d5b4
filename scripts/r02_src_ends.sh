#!/bin/bash
# source-level stall profile of the one-channel SIMT kernels (ncu --set full --import-source on), exported as CSV on the box
mkdir -p gpurun_out
export W=0 S=1
timeout 200 python scripts/profile_step.py > gpurun_out/r02_plain_step2.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_wgrad|stem_fwd|tail_fwd' -c 3 -o /tmp/ends2 -f python scripts/profile_step.py > gpurun_out/r02_ncu_ends2.log 2>&1
echo "rc=$?"
for k in stem_wgrad stem_fwd tail_fwd; do ncu -i /tmp/ends2.ncu-rep --page source --csv -k regex:$k > gpurun_out/r02_src_$k.csv 2>/dev/null; done
ncu -i /tmp/ends2.ncu-rep --page raw --csv > gpurun_out/r02_full_ends2.csv 2>/dev/null
ls -la gpurun_out/r02_src_*.csv
