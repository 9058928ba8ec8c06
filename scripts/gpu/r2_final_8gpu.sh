#!/bin/bash
# 8 GPUs, final build: the scaling line at N=8, BASELINE configs #3 (B=256 strong scaling) and #5 (256x256 K=16 T=8 B=64/GPU)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=8
timeout 600 $TR --nproc-per-node $N --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
timeout 600 $TR --nproc-per-node $N --master-port 29518 bench.py --gpus $N --config 3 --steps 10 --warmup 3 > gpurun_out/r2_cfg3_${N}gpu.json 2> gpurun_out/r2_cfg3_${N}gpu.err
timeout 900 $TR --nproc-per-node $N --master-port 29519 bench.py --gpus $N --config 5 --steps 3 --warmup 3 > gpurun_out/r2_cfg5_${N}gpu.json 2> gpurun_out/r2_cfg5_${N}gpu.err
for f in gpurun_out/r2_bench_8gpu gpurun_out/r2_cfg3_8gpu gpurun_out/r2_cfg5_8gpu; do echo == $f; tail -2 $f.err | cut -c1-200; cut -c1-330 $f.json; done
