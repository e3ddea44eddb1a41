timeout 300 python tools/solve_ab.py --batch 24 > gpurun_out/q3_ab.log 2>&1; tail -3 gpurun_out/q3_ab.log
for q in 0 2 8; do
  for st in 1 8; do
    AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_only.py --streams $st --steps 6 > gpurun_out/q3_quad_q${q}_st${st}.json 2>gpurun_out/q3_err.log; cat gpurun_out/q3_quad_q${q}_st${st}.json
  done
done
timeout 300 python tools/solve_only.py --streams 1 --batch 8192 --npts 10000 > gpurun_out/q3_quad_b8192.json 2>>gpurun_out/q3_err.log; cat gpurun_out/q3_quad_b8192.json
tail -5 gpurun_out/q3_err.log
