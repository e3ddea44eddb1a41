for q in 1 4; do AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_ab.py --batch 96 2>&1 | tail -1; done
for q in 0 1 2 4; do
  echo "Q=$q streams=1"; AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_only.py --streams 1 --steps 6 2>>gpurun_out/q7_err.log | tee gpurun_out/q7_quad_q${q}_st1.json
done
echo "Q=4 streams=8"; AMPC_QUADS_PER_WARP=4 timeout 300 python tools/solve_only.py --streams 8 --steps 6 2>>gpurun_out/q7_err.log | tee gpurun_out/q7_quad_q4_st8.json
echo "Q=4 B=32768"; AMPC_QUADS_PER_WARP=4 timeout 300 python tools/solve_only.py --streams 1 --steps 2 --batch 32768 --npts 4096 2>>gpurun_out/q7_err.log | tee gpurun_out/q7_quad_q4_b32768.json
AMPC_QUADS_PER_WARP=8 ncu --set full --clock-control none --import-source on -k regex:ipm_quad -s 1 -c 1 -o gpurun_out/q7_quad_q8_b16k python tools/solve_only.py --streams 1 --steps 1 --batch 16384 --npts 4096 > gpurun_out/q7_ncu.log 2>&1
AMPC_QUADS_PER_WARP=4 ncu --set full --clock-control none --import-source on -k regex:ipm_quad -s 1 -c 1 -o gpurun_out/q7_quad_q4_b16k python tools/solve_only.py --streams 1 --steps 1 --batch 16384 --npts 4096 >> gpurun_out/q7_ncu.log 2>&1
tail -3 gpurun_out/q7_ncu.log; tail -3 gpurun_out/q7_err.log
