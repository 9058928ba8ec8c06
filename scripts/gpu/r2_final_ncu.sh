#!/bin/bash
# ncu --set full of the decoder kernels of the committed build (kept apart from r2_final.sh: gpurun returns at most 64 MiB)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 4 -o gpurun_out/r2_final_tf32 python bench.py --precision tf32 --profile-mode --steps 1 --warmup 1 > gpurun_out/r2_final_tf32_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 4 -o gpurun_out/r2_final_fp16 python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > gpurun_out/r2_final_fp16_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
