for mode in fuse nofuse; do
  if [ $mode = nofuse ]; then export MMVAE_NO_BWD_FUSE=1; else unset MMVAE_NO_BWD_FUSE; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file gpurun_out/a6_launches_$mode.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/a6_ncu_$mode.log 2>&1
  echo "$mode rc=$? rows=$(grep -c gpu__time gpurun_out/a6_launches_$mode.csv)"
done
