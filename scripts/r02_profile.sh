#!/bin/bash
# r02 evidence pass on one B200: launch list of the bench command, ncu --set full of the pointwise / loss / 1-channel
# kernels and of the roofline launches (exported to CSV on the box: the .ncu-rep files exceed what gpurun copies back),
# stock-PyTorch baseline.  Each ncu run follows a plain run of the same command.
mkdir -p gpurun_out
O=gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --blocks 1 --no-extra --no-cpu-baseline"
full() {  # full <tag> <kernel regex> <skip> <count> <command...>
  tag=$1; rx=$2; skip=$3; cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o /tmp/$tag -f "$@" > $O/r02_ncu_$tag.log 2>&1
  rc=$?
  ncu -i /tmp/$tag.ncu-rep --page raw --csv > $O/r02_full_$tag.csv 2>/dev/null
  echo "full $tag rc=$rc kernels=$(($(wc -l < $O/r02_full_$tag.csv) - 2))"
}
if [[ " $* " == *" launches "* ]]; then
timeout 200 $BENCH > $O/r02_plain_bench.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv \
  --log-file $O/r02_step_launches.csv $BENCH > $O/r02_ncu_bench.log 2>&1
echo "launch list rc=$? lines=$(grep -c gpu__time $O/r02_step_launches.csv)"
fi
if [[ " $* " == *" full "* ]]; then
W=0 S=1 timeout 200 python scripts/profile_step.py > $O/r02_plain_step.log 2>&1 || exit 1
export W=0 S=1
full ends 'tail_bwd|tail_fwd|stem_fwd|stem_wgrad|loss_gauss|bn_apply_out|mmd_kernel|bn_bwd_c1|philox|heads_' 0 16 python scripts/profile_step.py

full bnapply 'bn_apply_kernel' 0 3 python scripts/profile_step.py
full bnbwd 'bn_bwd_sweep|bn_bwd_apply|bn_bwd_cluster|bn_bwd_reduce' 0 24 python scripts/profile_step.py
timeout 200 python scripts/profile_conv.py > $O/r02_plain_conv.log 2>&1 && \
full conv 'gconv_tc|slab_' 116 12 python scripts/profile_conv.py
fi
if [[ " $* " == *" torch "* ]]; then
timeout 300 python scripts/torch_gpu_baseline.py --steps 20 > $O/r02_torch_baseline.json 2> $O/r02_torch_baseline.err
echo "torch baseline rc=$?"
fi
du -sh $O; ls -la $O | head -30
