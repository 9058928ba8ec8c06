#!/bin/bash
# sweep: L2 prefetch distance of the row-streaming producers (IODINE_TC_PFD rows), tf32 and fp16
mkdir -p gpurun_out; : > gpurun_out/r2_pfd.txt
for prec in tf32 fp16; do for d in 0 8 16 32; do
IODINE_TC_PFD=$d python bench.py --precision $prec --steps 5 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$prec pfd=$d conv avg_launch_ms=%.4f TF=%.0f ms_per_step=%.3f value=%.0f'%(r['avg_launch_ms'], r['achieved'], l['ms_per_step'], l['value']))
" | tee -a gpurun_out/r2_pfd.txt
done; done
