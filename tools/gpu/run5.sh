timeout 300 python tools/solve_ab.py --batch 24 > gpurun_out/q6_ab.log 2>&1; tail -4 gpurun_out/q6_ab.log
timeout 300 python tools/solve_ab.py --batch 16 --N 30 --K 3 --warm cold > gpurun_out/q6_ab30.log 2>&1; tail -2 gpurun_out/q6_ab30.log
for q in 1 2 4 8; do
  for st in 1 8; do
    echo "Q=$q streams=$st"; AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_only.py --streams $st --steps 6 2>>gpurun_out/q6_err.log | tee gpurun_out/q6_quad_q${q}_st${st}.json
  done
done
for q in 2 4 8; do
  echo "Q=$q B=8192"; AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_only.py --streams 1 --steps 3 --batch 8192 --npts 10000 2>>gpurun_out/q6_err.log | tee gpurun_out/q6_quad_q${q}_b8192.json
done
echo "Q=4 B=32768"; AMPC_QUADS_PER_WARP=4 timeout 300 python tools/solve_only.py --streams 1 --steps 2 --batch 32768 --npts 4096 2>>gpurun_out/q6_err.log | tee gpurun_out/q6_quad_q4_b32768.json
echo "Q=8 B=32768"; AMPC_QUADS_PER_WARP=8 timeout 300 python tools/solve_only.py --streams 1 --steps 2 --batch 32768 --npts 4096 2>>gpurun_out/q6_err.log | tee gpurun_out/q6_quad_q8_b32768.json
tail -5 gpurun_out/q6_err.log
