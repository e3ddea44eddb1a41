timeout 600 python -m pytest tests/test_gpu_round.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/b1_bench.json 2> gpurun_out/b1_bench.err; tail -c 600 gpurun_out/b1_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/b1_bench.json'))
for k in ('value','ms_per_step','gpu_launches','stage_ms_per_step','single_stream','solver','clocks'):
    print(k, d.get(k))
print('e2e', {k:v for k,v in d['e2e'].items() if k!='cloud_upload'})
print('cloud_upload', d['e2e'].get('cloud_upload'))
r=d['roofline']; print('roofline', r['kernel'], r['achieved'], r['peak'], r['frac'], r['step_share'], r['cloud_index']['frac'], r['knn_stage']['frac'])
print('cold', d.get('cold_start')); print('rop', d.get('reference_operating_point')); print('cpu', d.get('cpu_baseline')); print('c0', d.get('cpu_c0'))
PY
timeout 600 python bench.py --mode best_of --scenes 1024 --steps 3 > gpurun_out/b1_bestof.json 2> gpurun_out/b1_bestof.err; tail -c 400 gpurun_out/b1_bestof.err; cat gpurun_out/b1_bestof.json | cut -c1-1500
timeout 600 python bench.py --mode scenes65536 --scenes 16384 --steps 3 > gpurun_out/b1_scenes.json 2> gpurun_out/b1_scenes.err; tail -c 400 gpurun_out/b1_scenes.err; cat gpurun_out/b1_scenes.json | cut -c1-800
