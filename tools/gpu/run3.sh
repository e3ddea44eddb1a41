ncu --set full --clock-control none --import-source on -k regex:ipm_quad -s 1 -c 1 -o gpurun_out/q4_quad_b1024 python tools/solve_only.py --streams 1 --steps 1 > gpurun_out/q4_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ipm_quad -s 1 -c 1 -o gpurun_out/q4_quad_b8192 python tools/solve_only.py --streams 1 --steps 1 --batch 8192 --npts 10000 >> gpurun_out/q4_ncu.log 2>&1
tail -5 gpurun_out/q4_ncu.log
