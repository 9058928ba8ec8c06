#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5 > gpurun_out/r2j_pytest.log
cat gpurun_out/r2j_pytest.log | cut -c1-600
for prec in tf32 fp16; do
timeout 600 python bench.py --precision $prec --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$prec value=%.0f ms_per_step=%.3f e2e=%.0f conv avg_launch_ms=%.4f'%(l['value'], l['ms_per_step'], l['e2e']['value'], r['avg_launch_ms']))
"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches_fp16.csv python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2j_launches_fp16.csv | head -14
