# two GPUs: in-library all-reduce tests + the 2-rank bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm.py -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
cat gpurun_out/bench_2gpu.json
