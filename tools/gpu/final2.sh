timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/f2.err; tail -c 300 gpurun_out/f2.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02.csv python tools/profile_step.py > gpurun_out/p1.log 2>&1; tail -1 gpurun_out/p1.log
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_step_r02 -f python tools/profile_step.py > gpurun_out/p2.log 2>&1; tail -1 gpurun_out/p2.log
python - <<'PY'
import json
def last(p): return json.loads([l for l in open(p) if l.startswith('{')][-1])
d=last('gpurun_out/bench_r02_1gpu.json')
print('value', d['value'], 'ms/step', d['ms_per_step'], d['stage_ms_per_step'], 'launches', d['gpu_launches'])
e=d['e2e']; print('e2e', e['value'], e['h2d_GBps'], 'u16', e['depth_u16']['value'], 'cloud_upload', e['cloud_upload']['value'])
r=d['roofline']; print('roofline', r['achieved'], r['peak'], r['frac'], r['avg_launch_ms'], r['per_launch']['frac'], r['in_step']['frac'], 'idx', r['cloud_index']['frac'], 'knn', r['knn_stage']['frac'])
print('single', d['single_stream']['value'], d['single_stream']['stage_ms'], 'cold', d['cold_start']['value'], 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])
PY
