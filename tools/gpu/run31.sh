N=$1
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/b31_$N.err
tail -c 300 gpurun_out/b31_$N.err
python - <<PY
import json
def last(p): return json.loads([l for l in open(p) if l.startswith('{')][-1])
d=last('gpurun_out/bench_r02_${N}gpu.json')
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['run']['per_rank'])
PY
