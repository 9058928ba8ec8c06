#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 600 python bench.py --precision $1 --steps 6 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r2aj_$1_$2.json 2> gpurun_out/r2aj_$1_$2.err
  echo "rc=$? $1 unit=$2"; tail -c 400 gpurun_out/r2aj_$1_$2.err; python -c "
import json,sys
t=open('gpurun_out/r2aj_$1_$2.json').read().strip()
if t:
    l=json.loads(t.splitlines()[-1]); print('$1 unit=$2', round(l['value']), round(l['ms_per_step'],3), round(l['roofline']['avg_launch_ms'],4))"
}
IODINE_TC_RS_UNIT=1 run tf32 1
IODINE_TC_RS_UNIT=2 run tf32 2
run tf32 default
IODINE_TC_RS_UNIT=2 run tf32 2b
IODINE_TC_RS_UNIT=7 run fp16 7
