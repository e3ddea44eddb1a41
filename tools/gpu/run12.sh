for q in 1 8; do AMPC_SOLVE_KERNEL=quad AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_ab.py --batch 96 2>&1 | tail -1; done
AMPC_SOLVE_KERNEL=quad timeout 300 python tools/solve_ab.py --batch 48 --N 30 --K 3 --warm cold 2>&1 | tail -1
for b in 8192 32768; do echo "quad B=$b"; timeout 300 python tools/solve_only.py --streams 1 --steps 3 --batch $b --npts 4096 2>>gpurun_out/q13_err.log | python -c "import json,sys; d=json.load(sys.stdin); print(d['stage_ms']['solve'], d['solves_per_s'], d['iters'], d['converged'])"; done
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_tick.py -x -q 2>&1 | tail -3
tail -3 gpurun_out/q13_err.log
