#!/bin/bash
mkdir -p gpurun_out
for prec in tf32 fp16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_${prec}_launches.csv python bench.py --precision $prec --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2_${prec}_launches.csv 2>/dev/null | head -8
done
