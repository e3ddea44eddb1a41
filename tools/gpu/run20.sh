for tool in racecheck synccheck; do
  echo "== compute-sanitizer $tool"
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_${tool}_r02.log python tools/solve_ab.py --batch 24 > gpurun_out/sanitizer_${tool}_r02.out 2>&1
  tail -2 gpurun_out/sanitizer_${tool}_r02.log; tail -1 gpurun_out/sanitizer_${tool}_r02.out
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
AMPC_SOLVE_KERNEL=warp timeout 300 python tools/solve_only.py --streams 1 --steps 6 | cut -c1-200
AMPC_SOLVE_KERNEL=warp timeout 300 python tools/solve_only.py --streams 8 --steps 6 | cut -c1-200
