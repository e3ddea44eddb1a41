timeout 300 python bench.py --mode knn_sweep --steps 5 > gpurun_out/knn_sweep_r02.json 2> gpurun_out/f3.err
timeout 300 python bench.py --mode best_of --steps 3 > gpurun_out/best_of_r02.json 2>> gpurun_out/f3.err
timeout 300 python bench.py --mode scenes65536 --steps 3 --chunk 32768 > gpurun_out/scenes65536_r02_1gpu.json 2>> gpurun_out/f3.err
python - <<'PY'
import json
def last(p): return json.loads([l for l in open(p) if l.startswith('{')][-1])
k=last('gpurun_out/knn_sweep_r02.json')
for r in k['rows']: print(r['npts'], round(r['index_ms'],3), round(r['search_ms'],3), round(r['stage_frac'],3))
for r in k['shuffled_storage_order']: print(r['npts'], r['layout'], round(r['index_ms'],2), round(r['search_ms'],2), round(r['stage_frac'],3))
b=last('gpurun_out/best_of_r02.json'); print('best_of', b['value'], b['ms_per_step'], b['argmin_parity_vs_host_reduction'])
s=last('gpurun_out/scenes65536_r02_1gpu.json'); print('scenes', s['value'], s['ms_per_job'])
PY
