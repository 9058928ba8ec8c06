# saved-activation L2 prefetch experiment: per-kernel ncu durations with and without it
mkdir -p gpurun_out
for d in 0 1; do
  if [ $d = 1 ]; then export IODINE_TC_NO_APF=1; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -c 160 --csv --log-file gpurun_out/apf_$d.csv python bench.py --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
  echo "== no_apf=$d"; python scripts/launch_summary.py gpurun_out/apf_$d.csv 2>/dev/null | grep conv_tc
done
