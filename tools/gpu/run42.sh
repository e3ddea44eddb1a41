AMPC_INDEX_SMALL=1 timeout 600 python -m pytest tests/test_gpu_knn.py -x -q 2>&1 | tail -1
run() { # name, env, args
  env $2 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $3 2> gpurun_out/b42.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$1 value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()})"
}
run "A default L2" "X=1" ""
run "B small  L2" "AMPC_INDEX_SMALL=1" ""
run "C small  L3" "AMPC_INDEX_SMALL=1" "--streams 3 --in-flight 96"
run "D small  L4" "AMPC_INDEX_SMALL=1" "--streams 4 --in-flight 128"
run "E default L3" "X=1" "--streams 3 --in-flight 96"
