#!/bin/bash
# 8 GPUs: 2-rank equality tests, the scaling line at N=8, BASELINE configs #3 (B=256 strong scaling) and #5 (256x256 K=16 T=8 B=64/GPU)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_multi_pytest.log
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ksplit.py -x -q -rs 2>&1 | tail -15 >> gpurun_out/r2_multi_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8; do
timeout 600 $TR --nproc-per-node $N --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
timeout 600 $TR --nproc-per-node $N --master-port 29518 bench.py --gpus $N --config 3 --steps 10 --warmup 3 > gpurun_out/r2_cfg3_${N}gpu.json 2> gpurun_out/r2_cfg3_${N}gpu.err
timeout 900 $TR --nproc-per-node $N --master-port 29519 bench.py --gpus $N --config 5 --steps 3 --warmup 3 > gpurun_out/r2_cfg5_${N}gpu.json 2> gpurun_out/r2_cfg5_${N}gpu.err
timeout 600 $TR --nproc-per-node $N --master-port 29520 bench.py --gpus $N --mode train --steps 5 --warmup 2 --precision fp16 > gpurun_out/r2_train_${N}gpu.json 2> gpurun_out/r2_train_${N}gpu.err
done
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r2_cfg3_1gpu.json 2> gpurun_out/r2_cfg3_1gpu.err
cat gpurun_out/r2_multi_pytest.log
for f in gpurun_out/r2_bench_8gpu gpurun_out/r2_cfg3_8gpu gpurun_out/r2_cfg5_8gpu gpurun_out/r2_train_8gpu gpurun_out/r2_cfg3_1gpu; do echo == $f; tail -2 $f.err | cut -c1-300; cut -c1-900 $f.json; done
