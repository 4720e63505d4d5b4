#!/bin/bash
# A/B of the library's experiment switches on the headline workload: one short bench per setting (ms per step).
# Usage: scripts/ab_env.sh TAG "VAR=val VAR2=val" "VAR=val" ...   ("" = baseline)
TAG=$1; shift
mkdir -p gpurun_out
for setting in "$@"; do
  out=$(env $setting timeout 120 python bench.py --no-extra --no-cpu-baseline --steps 100 --blocks 3 --warmup 10 2>/dev/null | \
        python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  e2e %.4f  adam %.4f  launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['with_adam']['ms_per_step'], d['launches_per_step']))")
  echo "[$setting] $out" | tee -a gpurun_out/${TAG}_ab.log
done
