#!/bin/bash
# rows per issue unit, now that the scout takes the waits (IODINE_TC_RS_UNIT; default tf32 2, fp16 5)
mkdir -p gpurun_out
run() {
  timeout 600 python bench.py --precision $1 --steps 6 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 unit=$2', round(l['value']), round(l['ms_per_step'],3), round(l['roofline']['avg_launch_ms'],4))"
}
for u in 1 2 3; do IODINE_TC_RS_UNIT=$u run tf32 $u; done
for u in 2 3 5 7; do IODINE_TC_RS_UNIT=$u run fp16 $u; done
