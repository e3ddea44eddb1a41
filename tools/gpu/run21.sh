ncu --set full --clock-control none --import-source on -k regex:ipm_quad -s 1 -c 1 -o gpurun_out/q22_quad_b64k python tools/solve_only.py --streams 1 --steps 1 --batch 65536 --npts 2048 > gpurun_out/q22_ncu.log 2>&1
tail -2 gpurun_out/q22_ncu.log
timeout 300 python tools/solve_only.py --streams 1 --steps 6 | cut -c1-220
