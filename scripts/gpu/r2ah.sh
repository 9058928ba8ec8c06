#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 tests/test_gpu_tc.py tests/test_gpu_at_size.py 2>&1 | tail -3
for prec in fp16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2ah_launches_$prec.csv python bench.py --precision $prec --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2ah_launches_$prec.csv 2>/dev/null | grep "refine\|kernel "
grep refine_tc gpurun_out/r2ah_launches_$prec.csv | tail -6 | awk -F'","' '{print $NF}'
done
