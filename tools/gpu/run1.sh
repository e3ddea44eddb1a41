timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q2_pytest.log 2>&1; tail -5 gpurun_out/q2_pytest.log
for q in 0 1 2 4 8; do
  for st in 1 8; do
    AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_only.py --streams $st --steps 6 > gpurun_out/q2_quad_q${q}_st${st}.json 2>gpurun_out/q2_err.log; cat gpurun_out/q2_quad_q${q}_st${st}.json
  done
done
AMPC_SOLVE_KERNEL=warp timeout 300 python tools/solve_only.py --streams 1 > gpurun_out/q2_warp_st1.json 2>>gpurun_out/q2_err.log; cat gpurun_out/q2_warp_st1.json
AMPC_SOLVE_KERNEL=warp timeout 300 python tools/solve_only.py --streams 8 > gpurun_out/q2_warp_st8.json 2>>gpurun_out/q2_err.log; cat gpurun_out/q2_warp_st8.json
timeout 300 python tools/solve_only.py --streams 1 --batch 8192 --npts 10000 > gpurun_out/q2_quad_b8192.json 2>>gpurun_out/q2_err.log; cat gpurun_out/q2_quad_b8192.json
tail -5 gpurun_out/q2_err.log
