#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q --timeout 600 2>&1 | tail -40 > gpurun_out/r2c_train.log
cat gpurun_out/r2c_train.log
