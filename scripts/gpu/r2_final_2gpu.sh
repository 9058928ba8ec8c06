#!/bin/bash
# two GPUs, final build: the 2-rank equality tests (all-reduce shards, K-split, training), the 2-rank bench line, K-split lines
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_two_gpu_pytest.log
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ksplit.py tests/test_gpu_train.py -x -q -rs 2>&1 | tail -6 >> gpurun_out/r2_two_gpu_pytest.log
cat gpurun_out/r2_two_gpu_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -2 gpurun_out/r2_bench_2gpu.err | cut -c1-300; cut -c1-400 gpurun_out/r2_bench_2gpu.json
for prec in tf32; do
timeout 600 python bench.py --batch 1 --slots 16 --steps 20 --warmup 3 --precision $prec --no-variants --no-cpu-baseline > gpurun_out/r2_ksplit_1gpu_$prec.json 2> gpurun_out/r2_ksplit_1gpu_$prec.err
timeout 600 $TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --k-split --batch 1 --slots 16 --steps 20 --warmup 3 --precision $prec --no-variants > gpurun_out/r2_ksplit_2gpu_$prec.json 2> gpurun_out/r2_ksplit_2gpu_$prec.err
for f in gpurun_out/r2_ksplit_1gpu_$prec gpurun_out/r2_ksplit_2gpu_$prec; do echo == $f; tail -2 $f.err | cut -c1-300; cut -c1-330 $f.json; done
done
