AMPC_SOLVE_KERNEL=quad timeout 300 python tools/solve_ab.py --batch 96 2>&1 | tail -1
for lib in avoid-mpc_b200/lib/variants/libampc_prev.so avoid-mpc_b200/lib/libampc.so; do for b in 65536; do echo "$lib B=$b"; AMPC_LIB=$PWD/$lib timeout 300 python tools/solve_only.py --streams 1 --steps 3 --batch $b --npts 2048 2>>gpurun_out/q25_err.log | python -c "import json,sys; d=json.load(sys.stdin); print(d['stage_ms'], d['solves_per_s'], d['iters'])"; done; done
tail -2 gpurun_out/q25_err.log
