for wps in 2 3 4 6 8; do
  echo "warps/SM $wps B=8192"; AMPC_QUAD_WARPS_PER_SM=$wps timeout 300 python tools/solve_only.py --streams 1 --steps 3 --batch 8192 --npts 10000 2>>gpurun_out/q5_err.log | tee gpurun_out/q5_w${wps}_b8192.json
done
for wps in 4 8; do
  echo "warps/SM $wps B=32768"; AMPC_QUAD_WARPS_PER_SM=$wps timeout 300 python tools/solve_only.py --streams 1 --steps 2 --batch 32768 --npts 4096 2>>gpurun_out/q5_err.log | tee gpurun_out/q5_w${wps}_b32768.json
done
tail -3 gpurun_out/q5_err.log
