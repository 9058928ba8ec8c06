#!/bin/bash
run() {
  timeout 300 python bench.py --precision $1 --steps 8 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(l['value']), round(l['ms_per_step'],3), round(l['roofline']['avg_launch_ms'],4))"
}
run tf32 default
IODINE_TC_NO_APF=1 run tf32 no_apf
IODINE_TC_LEAD=3 run tf32 lead3
run tf32 default_again
IODINE_TC_NO_APF=1 run fp16 no_apf
run fp16 default
