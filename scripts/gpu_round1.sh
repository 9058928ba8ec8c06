# Round artefacts on one B200: parity suite, bench lines, ncu launch list + full capture of the top kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_fp16.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_list.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_fp16.csv > gpurun_out/launches_fp16.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|mixture_kernel|refine_tc_kernel|tc_layer1_kernel|tc_class_sum_kernel" -s 20 -c 14 -o gpurun_out/prof_top python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
cat gpurun_out/bench_fp16.json
