#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5 > gpurun_out/r2i_pytest.log
cat gpurun_out/r2i_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print('tf32', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], 'conv avg ms', l['roofline']['avg_launch_ms'], 'clocks', l['clocks'])
print({k:(v['value']) for k,v in l.get('variants',{}).items()})
PY
