#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q --timeout 600 2>&1 | tail -15 > gpurun_out/r2g_train.log
cat gpurun_out/r2g_train.log | cut -c1-1500
for prec in fp16 tf32; do
timeout 600 python bench.py --mode train --steps 5 --warmup 2 --precision $prec > gpurun_out/r2g_train_$prec.json 2> gpurun_out/r2g_train_$prec.err
cut -c1-420 gpurun_out/r2g_train_$prec.json; tail -2 gpurun_out/r2g_train_$prec.err | cut -c1-300
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2g_train_launches.csv python bench.py --mode train --steps 1 --warmup 2 --precision fp16 > gpurun_out/r2g_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2g_train_launches.csv > gpurun_out/r2g_train_launches.md
head -32 gpurun_out/r2g_train_launches.md
