# usage: run30.sh N -- scaling trace: the default bench with the cost all-gather inline vs on its own stream
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
AMPC_BENCH_GATHER=inline timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/trace_inline_${N}gpu.json 2> gpurun_out/b30_$N.err
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2>> gpurun_out/b30_$N.err
tail -c 300 gpurun_out/b30_$N.err
python - <<PY
import json
def last(p): return json.loads([l for l in open(p) if l.startswith('{')][-1])
for f in ('trace_inline_${N}gpu','bench_r02_${N}gpu'):
    d=last('gpurun_out/%s.json'%f)
    print(f,'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['run']['collective'])
PY
