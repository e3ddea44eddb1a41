# usage: run27.sh N   -- default bench at N GPUs (driver's launch line)
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/b27_$N.err
tail -c 400 gpurun_out/b27_$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02_${N}gpu.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'e2e', d['e2e']['value'], d['e2e'].get('h2d_GBps'), 'clocks', d['clocks'])
PY
