ncu --set full --clock-control none --import-source on -k regex:ipm_quad -s 1 -c 1 -o gpurun_out/q12_quad_b32k python tools/solve_only.py --streams 1 --steps 1 --batch 32768 --npts 4096 > gpurun_out/q12_ncu.log 2>&1
tail -2 gpurun_out/q12_ncu.log
