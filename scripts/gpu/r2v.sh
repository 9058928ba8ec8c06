#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5 > gpurun_out/r2v_pytest.log
cat gpurun_out/r2v_pytest.log | cut -c1-800
run() {
  timeout 600 python bench.py --precision $1 --steps 6 --warmup 3 --no-cpu-baseline --no-variants 2>gpurun_out/r2v_$1_$2.err | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$1 $2 value=%.0f ms_per_step=%.3f e2e=%.0f conv avg_launch_ms=%.4f'%(l['value'], l['ms_per_step'], l['e2e']['value'], r['avg_launch_ms']))
"
}
run fp16 prod
run tf32 prod
IODINE_TC_DEBUG=26 run tf32 noinput_nostore_noact
IODINE_TC_DEBUG=26 run fp16 noinput_nostore_noact
IODINE_TC_DEBUG=27 run tf32 noinput_nostore_noact_notmem
IODINE_TC_DEBUG=27 run fp16 noinput_nostore_noact_notmem
