for q in 1 8; do AMPC_SOLVE_KERNEL=quad AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_ab.py --batch 96 2>&1 | tail -1; done
AMPC_SOLVE_KERNEL=quad timeout 300 python tools/solve_ab.py --batch 48 --N 30 --K 3 --warm cold 2>&1 | tail -1
for w in 4 5 6; do for b in 32768 65536; do echo "warps/SM=$w B=$b"; AMPC_QUAD_WARPS_PER_SM=$w timeout 300 python tools/solve_only.py --streams 1 --steps 3 --batch $b --npts 2048 2>>gpurun_out/q23_err.log | python -c "import json,sys; d=json.load(sys.stdin); print(d['stage_ms']['solve'], d['solves_per_s'])"; done; done
tail -2 gpurun_out/q23_err.log
