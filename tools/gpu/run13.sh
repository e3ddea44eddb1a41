for cfg in "1 32" "2 32" "4 32" "2 64"; do set -- $cfg; echo "lanes=$1 in_flight=$2"; timeout 900 python bench.py --steps 8 --warmup 3 --streams $1 --in-flight $2 --no-cpu-baseline > gpurun_out/b2_l$1_f$2.json 2> gpurun_out/b2.err; tail -c 300 gpurun_out/b2.err; python - <<PY
import json
d=json.load(open('gpurun_out/b2_l$1_f$2.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], d['stage_ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
PY
done
