timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/b4.err; tail -c 300 gpurun_out/b4.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r02_reference_arm.json 2>> gpurun_out/b4.err
timeout 900 python bench.py --mode best_of --steps 3 > gpurun_out/best_of_r02.json 2>> gpurun_out/b4.err
timeout 900 python bench.py --mode scenes65536 --steps 3 --chunk 32768 > gpurun_out/scenes65536_r02_1gpu.json 2>> gpurun_out/b4.err
timeout 900 python bench.py --mode knn_sweep --steps 5 > gpurun_out/knn_sweep_r02.json 2>> gpurun_out/b4.err
tail -c 600 gpurun_out/b4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_1gpu.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], d['stage_ms_per_step'], 'launches', d['gpu_launches'])
print('e2e', d['e2e']['value'], d['e2e']['h2d_GBps'], 'cloud_upload', d['e2e']['cloud_upload']['value'])
r=d['roofline']; print('roofline', r['kernel'], r['achieved'], r['peak'], r['frac'], r['traffic'], 'idx', r['cloud_index']['frac'], 'knn', r['knn_stage']['frac'])
print('single', d['single_stream']['value'], 'cold', d['cold_start']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('clocks', d['clocks'])
print('ref arm', json.load(open('gpurun_out/bench_r02_reference_arm.json'))['value'])
b=json.load(open('gpurun_out/best_of_r02.json')); print('best_of', b['value'], b['ms_per_step'], b['stage_ms'], b['argmin_parity_vs_host_reduction'], b['solver']['converged_frac'])
s=json.load(open('gpurun_out/scenes65536_r02_1gpu.json')); print('scenes', s['value'], s['ms_per_job'])
k=json.load(open('gpurun_out/knn_sweep_r02.json'))
for r in k['rows']: print(r['npts'], 'index %.3f search %.3f stage_frac %.3f idx_frac %.3f' % (r['index_ms'], r['search_ms'], r['stage_frac'], r['index_frac']))
for r in k['shuffled_storage_order']: print('shuffled', r['npts'], r['layout'], 'frac %.3f' % r['stage_frac'])
PY
