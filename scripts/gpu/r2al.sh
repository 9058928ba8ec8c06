#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
for prec in tf32 fp16; do
timeout 600 python bench.py --precision $prec --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$prec', round(l['value']), round(l['ms_per_step'],3), round(l['roofline']['avg_launch_ms'],4))"
IODINE_TC_DEBUG=26 timeout 600 python bench.py --precision $prec --steps 6 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$prec bits26', round(l['value']), round(l['ms_per_step'],3), round(l['roofline']['avg_launch_ms'],4))"
done
