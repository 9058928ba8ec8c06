#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q --timeout 600 2>&1 | tail -6 > gpurun_out/r2h_train.log
cat gpurun_out/r2h_train.log | cut -c1-1500
for prec in tf32 fp16; do
timeout 600 python bench.py --mode train --steps 5 --warmup 2 --precision $prec > gpurun_out/r2h_train_$prec.json 2> gpurun_out/r2h_train_$prec.err
cut -c1-420 gpurun_out/r2h_train_$prec.json; tail -2 gpurun_out/r2h_train_$prec.err | cut -c1-300
done
