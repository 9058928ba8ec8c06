#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
for prec in tf32 fp16; do
timeout 600 python bench.py --precision $prec --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$prec', round(l['value']), round(l['ms_per_step'],3), round(l['roofline']['avg_launch_ms'],4))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2ak_launches_fp16.csv python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2ak_launches_fp16.csv 2>/dev/null | head -14
