timeout 900 python -m pytest tests/test_gpu_knn.py -x -q 2>&1 | tail -2
timeout 600 python bench.py --mode knn_sweep --steps 5 > gpurun_out/knn_sweep_try.json 2> gpurun_out/b44.err; tail -c 300 gpurun_out/b44.err
python - <<'PY'
import json
k=json.loads([l for l in open('gpurun_out/knn_sweep_try.json') if l.startswith('{')][-1])
for r in k['shuffled_storage_order']: print('shuffled', r['npts'], r['layout'], 'index %.2f search %.2f frac %.3f' % (r['index_ms'], r['search_ms'], r['stage_frac']), r.get('identical_to_unorganised'))
PY
