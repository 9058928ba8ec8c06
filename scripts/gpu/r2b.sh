#!/bin/bash
# round 2, pass b: default bench line (tf32 = configs[1]) with variants + CPU arm, ncu launch lists (tf32, fp16)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_ref.json 2> gpurun_out/r2b_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches_tf32.csv python bench.py --precision tf32 --profile-mode --steps 1 --warmup 1 > gpurun_out/r2b_ncu_tf32.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches_fp16.csv python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > gpurun_out/r2b_ncu_fp16.log 2>&1
python scripts/launch_summary.py gpurun_out/r2b_launches_tf32.csv > gpurun_out/r2b_launches_tf32.md
python scripts/launch_summary.py gpurun_out/r2b_launches_fp16.csv > gpurun_out/r2b_launches_fp16.md
cat gpurun_out/r2b_bench.json gpurun_out/r2b_ref.json; tail -3 gpurun_out/r2b_bench.err
