#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --mode train --steps 5 --warmup 2 --precision fp16 > gpurun_out/r2d_train_fp16.json 2> gpurun_out/r2d_train_fp16.err
timeout 900 python bench.py --mode train --steps 5 --warmup 2 --precision tf32 > gpurun_out/r2d_train_tf32.json 2> gpurun_out/r2d_train_tf32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2d_train_launches.csv python bench.py --mode train --steps 1 --warmup 2 --precision fp16 > gpurun_out/r2d_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2d_train_launches.csv > gpurun_out/r2d_train_launches.md
cat gpurun_out/r2d_train_fp16.json gpurun_out/r2d_train_tf32.json; tail -3 gpurun_out/r2d_train_fp16.err; head -30 gpurun_out/r2d_train_launches.md
