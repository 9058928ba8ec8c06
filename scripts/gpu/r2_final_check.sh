#!/bin/bash
# last regression pass of the round: every GPU test, smoke(), the remaining single-GPU BASELINE configurations, training lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4 > gpurun_out/r2_final_pytest.log
cat gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
for cfg in 1 4 5; do
timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r2_cfg$cfg.json 2> gpurun_out/r2_cfg$cfg.err
python -c "
import json
l=json.loads(open('gpurun_out/r2_cfg$cfg.json').read().strip().splitlines()[-1]); print('cfg$cfg', round(l['value']), round(l['ms_per_step'],3), 'e2e', round(l['e2e']['value']), l['config']['precision'])"
done
for prec in fp16 tf32; do
timeout 600 python bench.py --mode train --precision $prec --steps 5 --warmup 2 > gpurun_out/r2_train_${prec}_1gpu.json 2> gpurun_out/r2_train_${prec}_1gpu.err
cut -c1-260 gpurun_out/r2_train_${prec}_1gpu.json
done
