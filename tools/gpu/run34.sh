timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_round.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --mode knn_sweep --steps 5 > gpurun_out/knn_sweep_try.json 2> gpurun_out/b34.err
python - <<'PY'
import json
k=json.loads([l for l in open('gpurun_out/knn_sweep_try.json') if l.startswith('{')][-1])
for r in k['rows']: print(r['npts'], 'index %.3f search %.3f stage_frac %.3f' % (r['index_ms'], r['search_ms'], r['stage_frac']))
for r in k['shuffled_storage_order']: print('shuffled', r['npts'], r['layout'], 'search %.3f frac %.3f' % (r['search_ms'], r['stage_frac']))
PY
