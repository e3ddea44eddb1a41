timeout 600 python -m pytest tests/test_gpu_solve.py -x -q -k "refill or matches_oracle" 2>&1 | tail -2
for W in 1 2 3 6; do
  AMPC_QUAD_CTA_WARPS=$W timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/b39.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
r=d['roofline']
print('CTA warps=$W value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'solve-only tflops', round(r['achieved'],3), 'launch ms', round(r['avg_launch_ms'],2))"
done
