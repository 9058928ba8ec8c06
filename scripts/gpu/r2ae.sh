#!/bin/bash
mkdir -p gpurun_out
for prec in fp16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2ae_launches_$prec.csv python bench.py --precision $prec --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2ae_launches_$prec.csv 2>/dev/null | grep "layer1\|sample_u\|kernel "
done
timeout 600 ncu --set full --clock-control none -k regex:"tc_layer1" -s 4 -c 1 -o gpurun_out/r2ae_l1 python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/
