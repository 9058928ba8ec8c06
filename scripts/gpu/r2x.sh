#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5 > gpurun_out/r2x_pytest.log
cat gpurun_out/r2x_pytest.log | cut -c1-800
run() {
  timeout 600 python bench.py --precision $1 --steps 6 --warmup 3 --no-cpu-baseline --no-variants 2>gpurun_out/r2x_$1_$2.err | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$1 $2 value=%.0f ms_per_step=%.3f e2e=%.0f conv avg_launch_ms=%.4f'%(l['value'], l['ms_per_step'], l['e2e']['value'], r['avg_launch_ms']))
"
}
run fp16 prod
run tf32 prod
run fp16 prod_again
run tf32 prod_again
IODINE_TC_VERBOSE=1 timeout 300 python bench.py --precision fp16 --steps 1 --warmup 1 --no-cpu-baseline --no-variants 2>&1 | grep -i "ring\|geom\|R=" | head -5
for prec in fp16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches_$prec.csv python bench.py --precision $prec --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2x_launches_$prec.csv 2>/dev/null | head -9
done
