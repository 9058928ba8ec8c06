mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu --format=csv,noheader -lms 50 > gpurun_out/clocks_trace.csv &
SMI=$!
sleep 1
python bench.py --precision fp16 --steps 150 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('value=%.0f ms_per_step=%.2f conv avg_launch_ms=%.4f TF=%.0f clocks=%s'%(l['value'], l['ms_per_step'], r['avg_launch_ms'], r['achieved'], l['clocks']))"
kill $SMI
python - <<'PY'
rows=[l.strip().split(', ') for l in open('gpurun_out/clocks_trace.csv') if l.strip()]
sm=[int(r[0].split()[0]) for r in rows]; pw=[float(r[2].split()[0]) for r in rows]
print('samples',len(rows),'sm MHz min/median/max',min(sm),sorted(sm)[len(sm)//2],max(sm),'power max',max(pw))
busy=[(s,p,r[3]) for s,p,r in zip(sm,pw,rows) if p>500]
print('under load (>500W):',len(busy),'samples; sm MHz median',sorted(b[0] for b in busy)[len(busy)//2] if busy else None, 'power median', sorted(b[1] for b in busy)[len(busy)//2] if busy else None, 'sw_power_cap', set(b[2] for b in busy))
PY
