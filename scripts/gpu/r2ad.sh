#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8 > gpurun_out/r2ad_pytest.log
cat gpurun_out/r2ad_pytest.log | cut -c1-1200
for prec in tf32 fp16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2ad_launches_$prec.csv python bench.py --precision $prec --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2ad_launches_$prec.csv 2>/dev/null | grep "layer1\|sample_u\|kernel "
timeout 600 python bench.py --precision $prec --steps 6 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$prec', l['value'], l['ms_per_step'], l['roofline']['avg_launch_ms'])"
done
