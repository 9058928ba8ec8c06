#!/bin/bash
# two GPUs: the in-library all-reduce equality tests (2 ranks vs single process) + a 2-rank bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_comm2_pytest.log
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ksplit.py -x -q -rs 2>&1 | tail -15 >> gpurun_out/r2_comm2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
cat gpurun_out/r2_comm2_pytest.log; tail -3 gpurun_out/r2_bench_2gpu.err; cat gpurun_out/r2_bench_2gpu.json
