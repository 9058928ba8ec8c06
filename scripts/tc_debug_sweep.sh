# timing experiments: which part of conv_tc bounds the MMA rate (IODINE_TC_DEBUG bits, see conv_tc.cu)
mkdir -p gpurun_out
for d in ${SWEEP:-0 1 2 3 8 11}; do
  IODINE_TC_DEBUG=$d python bench.py --precision fp16 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('dbg=$d conv avg_launch_ms=%.4f TF=%.0f ms_per_step=%.2f'%(r['avg_launch_ms'], r['achieved'], l['ms_per_step']))
" | tee -a gpurun_out/tc_debug_sweep.txt
done
