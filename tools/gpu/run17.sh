timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "2 64" "1 64" "2 32" "4 64"; do set -- $cfg; echo "lanes=$1 in_flight=$2"; timeout 900 python bench.py --steps 6 --warmup 3 --streams $1 --in-flight $2 --no-cpu-baseline > gpurun_out/b3_l$1_f$2.json 2> gpurun_out/b3.err; tail -c 300 gpurun_out/b3.err; python - <<PY
import json
d=json.load(open('gpurun_out/b3_l$1_f$2.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], d['stage_ms_per_step'], 'frac', d['roofline']['frac'], 'single', d['single_stream']['value'])
PY
done
