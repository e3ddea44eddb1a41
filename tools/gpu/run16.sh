AMPC_SOLVE_KERNEL=quad AMPC_QUADS_PER_WARP=2 AMPC_QUAD_WARPS_PER_SM=1 timeout 300 python tools/solve_ab.py --batch 400 2>&1 | tail -1
for o in 0 1; do for b in 16384 32768 65536; do echo "order=$o B=$b"; AMPC_QUAD_ORDER=$o timeout 300 python tools/solve_only.py --streams 1 --steps 3 --batch $b --npts 2048 2>>gpurun_out/q16_err.log | python -c "import json,sys; d=json.load(sys.stdin); print(d['stage_ms']['solve'], d['solves_per_s'])"; done; done
tail -2 gpurun_out/q16_err.log
