#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_at_size.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 2>&1 | tail -3
for prec in fp16 tf32; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2an_launches_$prec.csv python bench.py --precision $prec --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2an_launches_$prec.csv 2>/dev/null | grep "layer1"
done
