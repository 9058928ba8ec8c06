#!/bin/bash
# ncu --set full of the tail kernels (fp16 mode)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mixture_kernel|assemble16|refine_tc_kernel|tc_layer1|post_grads|sample_u|linear_kernel" -s 20 -c 14 -o gpurun_out/r2k_tail python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > gpurun_out/r2k_ncu.log 2>&1
tail -3 gpurun_out/r2k_ncu.log | cut -c1-200
ls -la gpurun_out/r2k_tail.ncu-rep
