# sanity lines for the other BASELINE configurations (one GPU)
mkdir -p gpurun_out
python bench.py --precision bf16 --batch 256 --steps 2 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null > gpurun_out/cfg3.json
python bench.py --slots 11 --iters 7 --steps 3 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null > gpurun_out/cfg4.json
for f in cfg3 cfg4; do python -c "
import json,sys
l=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', round(l['value']), round(l['e2e']['value']), round(l['ms_per_step'],1))"; done
