timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/f1b.err; tail -c 300 gpurun_out/f1b.err
timeout 900 python bench.py --mode knn_sweep --steps 5 > gpurun_out/knn_sweep_r02.json 2>> gpurun_out/f1b.err
timeout 900 python bench.py --mode best_of --steps 3 > gpurun_out/best_of_r02.json 2>> gpurun_out/f1b.err
python - <<'PY'
import json
def last(p): return json.loads([l for l in open(p) if l.startswith('{')][-1])
d=last('gpurun_out/bench_r02_1gpu.json')
print('value', d['value'], 'ms/step', d['ms_per_step'], d['stage_ms_per_step'])
e=d['e2e']; print('e2e', e['value'], e['h2d_GBps'], 'u16', e['depth_u16']['value'])
r=d['roofline']; print('roofline', r['frac'], 'idx', r['cloud_index']['frac'], r['cloud_index']['traffic'], 'knn', r['knn_stage']['frac'], r['knn_stage']['search_ms'])
print('single', d['single_stream'])
k=last('gpurun_out/knn_sweep_r02.json')
for r in k['rows']: print(r['npts'], round(r['index_ms'],3), round(r['search_ms'],3), round(r['stage_frac'],3))
for r in k['shuffled_storage_order']: print(r['npts'], r['layout'], round(r['index_ms'],2), round(r['search_ms'],2), round(r['stage_frac'],3))
b=last('gpurun_out/best_of_r02.json'); print('best_of', b['value'], b['ms_per_step'], b['argmin_parity_vs_host_reduction'])
PY
