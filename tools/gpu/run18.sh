timeout 1200 python -m pytest tests/test_gpu_knn.py tests/test_gpu_round.py tests/test_gpu_shim.py tests/test_gpu_depth.py -x -q 2>&1 | tail -3
for t in 0 1; do echo "two_level=$t"; AMPC_KNN_TWO_LEVEL=$t timeout 600 python bench.py --mode knn_sweep --steps 5 > gpurun_out/k1_sweep_$t.json 2>gpurun_out/k1.err; tail -c 300 gpurun_out/k1.err; python - <<PY
import json
d=json.load(open('gpurun_out/k1_sweep_$t.json'))
for r in d['rows']: print(r['npts'], 'index %.3f search %.3f stage_frac %.3f' % (r['index_ms'], r['search_ms'], r['stage_frac']))
for r in d['shuffled_storage_order']: print('shuffled', r['npts'], r['layout'], 'index %.3f search %.3f frac %.3f' % (r['index_ms'], r['search_ms'], r['stage_frac']), r.get('identical_to_unorganised'))
PY
done
