python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size)
import ctypes
rt=ctypes.CDLL('libcudart.so.12')
v=ctypes.c_int()
for name,attr in (('MaxAccessPolicyWindowSize',109),('MaxPersistingL2CacheSize',108)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0); print(name, v.value)
PY
for Pct in 0 40 60 80; do
  AMPC_QUAD_L2_PERSIST=$Pct timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/b43.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
r=d['roofline']
print('persist=$Pct value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'solve-only tflops', round(r['achieved'],3), 'launch ms', round(r['avg_launch_ms'],2))"
  tail -c 200 gpurun_out/b43.err
done
