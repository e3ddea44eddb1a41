timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/b29.err; tail -c 600 gpurun_out/b29.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02_1gpu.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], d['stage_ms_per_step'], 'launches', d['gpu_launches'])
e=d['e2e']; print('e2e', e['value'], e['h2d_GBps'], 'u16', e['depth_u16']['value'], e['depth_u16']['h2d_GBps'], 'cloud_upload', e['cloud_upload']['value'], e['host_binding'])
r=d['roofline']; print('roofline', r['kernel'], r['achieved'], r['peak'], r['frac'], r['traffic'], 'idx', r['cloud_index']['frac'], r['cloud_index']['avg_launch_ms'], 'knn', r['knn_stage']['frac'])
print('single', d['single_stream']['value'], 'cold', d['cold_start']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
