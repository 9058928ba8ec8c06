#!/bin/bash
# round-2 record of the committed build: smoke, default bench line (tf32 = configs[1]), fp16 variant line, reference arm,
# launch lists, and ncu --set full of the aux-path kernels (fp16 run) and of the tf32 decoder kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --precision fp16 --steps 20 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r2_bench_fp16.json 2> gpurun_out/r2_bench_fp16.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
for prec in tf32 fp16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_${prec}_launches.csv python bench.py --precision $prec --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mixture_fast|refine_l0f|refine_tc_kernel|tc_layer1" -s 10 -c 6 -o gpurun_out/r2_final_aux python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > gpurun_out/r2_final_aux_ncu.log 2>&1
cut -c1-600 gpurun_out/r2_bench_default.json; echo; cut -c1-300 gpurun_out/r2_bench_fp16.json; echo; cut -c1-400 gpurun_out/r2_bench_reference_arm.json
ls -la gpurun_out/*.ncu-rep | tail -4
