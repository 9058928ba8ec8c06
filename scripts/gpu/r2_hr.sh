#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_tc.py -x -q --timeout 600 -k "tf32 or tc" 2>&1 | tail -8 > gpurun_out/r2_hr_pytest.log
cat gpurun_out/r2_hr_pytest.log | cut -c1-800
: > gpurun_out/r2_hr.txt
for u in 2 1 3 4; do
IODINE_TC_RS_UNIT=$u python bench.py --precision tf32 --steps 5 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('tf32 half-row ring rs_unit=$u conv avg_launch_ms=%.4f TF=%.0f ms_per_step=%.3f value=%.0f'%(r['avg_launch_ms'], r['achieved'], l['ms_per_step'], l['value']))
" | tee -a gpurun_out/r2_hr.txt
done
