#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_at_size.py 2>&1 | tail -5 > gpurun_out/r2p_pytest.log
cat gpurun_out/r2p_pytest.log | cut -c1-800
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches_fp16.csv python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2p_launches_fp16.csv 2>/dev/null | head -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"refine_l0f|mixture" -s 4 -c 2 -o gpurun_out/r2p_l0f python bench.py --precision fp16 --profile-mode --steps 1 --warmup 1 > gpurun_out/r2p_ncu.log 2>&1
