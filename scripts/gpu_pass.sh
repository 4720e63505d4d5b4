#!/bin/bash
# One gpurun call on one B200: parity tests, bench line, conv timeline, ncu launch list of the bench command.
# Usage: scripts/gpu_pass.sh TAG [tests] [bench] [trace] [launches] [full:<kernel-regex>:<skip>:<count>]
TAG=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests)    timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${TAG}_tests.log ;;
    bench)    timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json ;;
    trace)    timeout 300 python scripts/trace_conv.py > gpurun_out/${TAG}_trace.log 2>&1; echo "trace rc=$?"; cat gpurun_out/${TAG}_trace.log ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv \
                --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
              echo "launches rc=$?"; grep -c gpu__time gpurun_out/${TAG}_launches.csv ;;
    full:*)   IFS=: read -r _ rx skip cnt <<< "$what"
              timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s "$skip" -c "$cnt" \
                -o gpurun_out/${TAG}_full_${rx//[^a-zA-Z0-9]/_} -f python scripts/profile_step.py > gpurun_out/${TAG}_ncu_full.log 2>&1
              echo "full rc=$?"; tail -n 3 gpurun_out/${TAG}_ncu_full.log ;;
  esac
done
