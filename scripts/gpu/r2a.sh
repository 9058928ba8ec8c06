#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-variants > gpurun_out/r2a_bench_fp16.json 2> gpurun_out/r2a_bench_fp16.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-variants --no-cpu-baseline --precision tf32 > gpurun_out/r2a_bench_tf32.json 2> gpurun_out/r2a_bench_tf32.err
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench_fp16.json gpurun_out/r2a_bench_tf32.json; tail -5 gpurun_out/r2a_bench_tf32.err
