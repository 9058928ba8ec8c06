set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
python bench.py --precision fp16 --steps 10 --warmup 3 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_fp16.csv python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 20 -c 4 -o gpurun_out/prof_conv_tc python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_fp16.json
