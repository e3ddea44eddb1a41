for F in 64 96 128; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --in-flight $F 2> gpurun_out/b35.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('F=$F value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), d['stage_ms_per_step'])"
  tail -c 200 gpurun_out/b35.err
done
