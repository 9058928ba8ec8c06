#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q --timeout 600 -k "clevr6_size" -s 2>&1 | tail -40 > gpurun_out/r2f_train.log
cat gpurun_out/r2f_train.log | cut -c1-2500
timeout 600 python bench.py --mode train --steps 5 --warmup 2 --precision fp16 > gpurun_out/r2f_train_fp16.json 2> gpurun_out/r2f_train_fp16.err
cut -c1-700 gpurun_out/r2f_train_fp16.json; tail -3 gpurun_out/r2f_train_fp16.err
