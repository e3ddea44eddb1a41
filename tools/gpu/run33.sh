for W in 6 5 4; do
  AMPC_QUAD_WARPS_PER_SM=$W timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/b33.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('W=$W value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), d['stage_ms_per_step'])"
done
for S in 3 4; do
  AMPC_QUAD_WARPS_PER_SM=5 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --streams $S --in-flight $((S*32)) 2> gpurun_out/b33.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('W=5 streams=$S value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), d['stage_ms_per_step'])"
done
