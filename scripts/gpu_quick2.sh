# quick GPU check: all GPU tests, fp16 bench line, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --precision fp16 --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | tee gpurun_out/bench_quick.json | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('value=%.0f e2e=%.0f ms_per_step=%.2f conv avg_launch_ms=%.4f TF=%.0f share=%.2f'%(l['value'], l['e2e']['value'], l['ms_per_step'], r['avg_launch_ms'], r['achieved'], r['share_of_step']))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_q.csv python bench.py --profile-mode --steps 1 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/launches_q.csv | head -14
