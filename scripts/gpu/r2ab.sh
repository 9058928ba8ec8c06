#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5 > gpurun_out/r2ab_pytest.log
cat gpurun_out/r2ab_pytest.log | cut -c1-800
bash scripts/gpu/r2_sanitize.sh 2>&1 | grep -E "==|SUMMARY"
for prec in tf32 fp16; do
timeout 600 python bench.py --precision $prec --steps 6 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$prec', l['value'], l['ms_per_step'], l['roofline']['avg_launch_ms'], l.get('aux_fuse',{}).get('frac'))"
done
