# per-kernel ncu durations under the IODINE_TC_DEBUG bits (1 no TMEM reads, 2 no epilogue stores, 4 no TMA, 8 no activation loads)
mkdir -p gpurun_out
for d in ${SWEEP:-0 2 8 10 11 15}; do
  IODINE_TC_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -c 160 --csv --log-file gpurun_out/dbg_$d.csv python bench.py --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
  echo "== dbg=$d"; python scripts/launch_summary.py gpurun_out/dbg_$d.csv 2>/dev/null | grep conv_tc
done
