N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/f46.err; echo rc=$?
grep -c . gpurun_out/bench_r02_${N}gpu.json; head -c 60 gpurun_out/bench_r02_${N}gpu.json; echo
timeout 300 $TR bench.py --gpus $N --mode scenes65536 --steps 3 2>> gpurun_out/f46.err | head -c 200; echo
