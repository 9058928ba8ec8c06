# producer run-ahead cap experiment: per-kernel ncu durations for IODINE_TC_LEAD values
mkdir -p gpurun_out
for d in ${SWEEP:-0 3 5 7}; do
  IODINE_TC_LEAD=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -c 160 --csv --log-file gpurun_out/lead_$d.csv python bench.py --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
  echo "== lead=$d"; python scripts/launch_summary.py gpurun_out/lead_$d.csv 2>/dev/null | grep conv_tc
done
