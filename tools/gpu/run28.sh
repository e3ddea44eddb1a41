# usage: run28.sh N   -- default bench + the 65536-scene job at N GPUs (driver's launch line)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/b28_$N.err
timeout 600 $TR bench.py --gpus $N --mode scenes65536 --steps 3 > gpurun_out/scenes65536_r02_${N}gpu.json 2>> gpurun_out/b28_$N.err
tail -c 400 gpurun_out/b28_$N.err
python - <<PY
import json
def last(p):
    return json.loads([l for l in open(p) if l.startswith('{')][-1])
d=last('gpurun_out/bench_r02_${N}gpu.json')
print('value', d['value'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'e2e', d['e2e']['value'], d['e2e'].get('h2d_GBps'), 'clocks', d['clocks'])
s=last('gpurun_out/scenes65536_r02_${N}gpu.json'); print('scenes', s['value'], s.get('ms_per_job'))
PY
