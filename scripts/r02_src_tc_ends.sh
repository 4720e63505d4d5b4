#!/bin/bash
# source-level stall profile of the tensor-core stem / tail kernels (ncu --set full --import-source on), exported as CSV on the box
mkdir -p gpurun_out
export W=0 S=1
timeout 200 python scripts/profile_step.py > gpurun_out/r02_plain_step3.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_.*_tc|tail_.*_tc' -c 4 -o /tmp/ends3 -f python scripts/profile_step.py > gpurun_out/r02_ncu_ends3.log 2>&1
echo "rc=$?"
for k in stem_wgrad_tc stem_fwd_tc tail_fwd_tc tail_bwd_tc; do ncu -i /tmp/ends3.ncu-rep --page source --csv -k regex:$k > gpurun_out/r02_src_$k.csv 2>/dev/null; done
ncu -i /tmp/ends3.ncu-rep --page raw --csv > gpurun_out/r02_full_tc_ends.csv 2>/dev/null
ls -la gpurun_out/r02_src_*_tc.csv
