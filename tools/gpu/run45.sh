N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_final_${N}gpu.json 2> gpurun_out/f45.err; echo rc=$?
tail -c 300 gpurun_out/f45.err
timeout 300 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 3 > gpurun_out/ref_final_${N}gpu.json 2>> gpurun_out/f45.err; echo rc=$?
python - <<PY
import json
def last(p): return json.loads([l for l in open(p) if l.startswith('{')][-1])
d=last('gpurun_out/bench_final_2gpu.json'); e=d['e2e']; r=d['roofline']
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', e['value'], 'roofline', r['frac'], r['kernel'], d['run']['per_rank'])
print(len([l for l in open('gpurun_out/bench_final_2gpu.json') if l.strip()]), 'stdout lines')
x=last('gpurun_out/ref_final_2gpu.json'); print('ref', x['value'], x['n_gpus'], x['config']==d['config'])
PY
