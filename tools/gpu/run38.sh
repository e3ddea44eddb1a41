timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/f38.err; tail -c 400 gpurun_out/f38.err
python - <<'PY'
import json
def last(p): return json.loads([l for l in open(p) if l.startswith('{')][-1])
d=last('gpurun_out/bench_r02_1gpu.json')
print('value', d['value'], 'ms/step', d['ms_per_step'], d['stage_ms_per_step'])
r=d['roofline']; print({k:r[k] for k in ('achieved','peak','frac','avg_launch_ms','launches_in_flight','iterations_per_launch','per_launch','in_step','step_share','traffic')})
PY
