# usage: finalN.sh N [sweep] -- final multi-GPU artefacts
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/fN_$N.err
timeout 600 $TR bench.py --gpus $N --mode scenes65536 --steps 3 > gpurun_out/scenes65536_r02_${N}gpu.json 2>> gpurun_out/fN_$N.err
if [ "$2" = "sweep" ]; then timeout 600 $TR bench.py --gpus $N --mode knn_sweep --steps 5 > gpurun_out/knn_sweep_r02_${N}gpu.json 2>> gpurun_out/fN_$N.err; fi
tail -c 300 gpurun_out/fN_$N.err
python - <<PY
import json, os
def last(p): return json.loads([l for l in open(p) if l.startswith('{')][-1])
d=last('gpurun_out/bench_r02_${N}gpu.json'); e=d['e2e']
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', e['value'], e['h2d_GBps'], 'u16', e['depth_u16']['value'], 'alone', e['pinned_h2d_GBps_alone'])
print(d['run']['per_rank'])
s=last('gpurun_out/scenes65536_r02_${N}gpu.json'); print('scenes', s['value'], s.get('ms_per_job'))
p='gpurun_out/knn_sweep_r02_${N}gpu.json'
if os.path.exists(p):
    k=last(p); print('sweep', ' '.join('%d:%.3f' % (r['npts'], r['stage_frac']) for r in k['rows']), k.get('n_gpus'))
PY
