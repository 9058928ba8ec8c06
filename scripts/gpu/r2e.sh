#!/bin/bash
# one GPU: BASELINE configs #1 and #4 bench lines, full GPU test-suite, ncu --set full of the top tf32 kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8 > gpurun_out/r2e_pytest.log
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 --no-variants > gpurun_out/r2e_cfg4.json 2> gpurun_out/r2e_cfg4.err
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 --no-variants > gpurun_out/r2e_cfg1.json 2> gpurun_out/r2e_cfg1.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2e_cfg5_1gpu.json 2> gpurun_out/r2e_cfg5_1gpu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 6 -o gpurun_out/r2e_prof_tf32 python bench.py --precision tf32 --profile-mode --steps 1 --warmup 1 > gpurun_out/r2e_ncu_full.log 2>&1
cat gpurun_out/r2e_pytest.log
for f in gpurun_out/r2e_cfg4 gpurun_out/r2e_cfg1 gpurun_out/r2e_cfg5_1gpu; do echo == $f; tail -2 $f.err | cut -c1-300; cut -c1-1200 $f.json; done
tail -5 gpurun_out/r2e_ncu_full.log
