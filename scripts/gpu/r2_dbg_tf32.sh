#!/bin/bash
# timing experiments: which part of the tf32 conv_tc kernels bounds them (IODINE_TC_DEBUG bits; results are wrong when set)
mkdir -p gpurun_out; : > gpurun_out/r2_dbg_tf32.txt
for d in 0 2 8 10 11; do
IODINE_TC_DEBUG=$d python bench.py --precision tf32 --steps 3 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('tf32 dbg=$d conv avg_launch_ms=%.4f TF=%.0f ms_per_step=%.3f'%(r['avg_launch_ms'], r['achieved'], l['ms_per_step']))
" | tee -a gpurun_out/r2_dbg_tf32.txt
done
for u in 1 2 3; do
IODINE_TC_RS_UNIT=$u python bench.py --precision tf32 --steps 3 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('tf32 rs_unit=$u conv avg_launch_ms=%.4f TF=%.0f ms_per_step=%.3f'%(r['avg_launch_ms'], r['achieved'], l['ms_per_step']))
" | tee -a gpurun_out/r2_dbg_tf32.txt
done
