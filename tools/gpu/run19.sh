tools/bin/dmma_condensed 16384 4 | tee gpurun_out/dmma_condensed_r02.json
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer $tool"
  AMPC_SOLVE_KERNEL=quad AMPC_QUADS_PER_WARP=8 timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_${tool}_r02.log python tools/solve_ab.py --batch 24 > gpurun_out/sanitizer_${tool}_r02.out 2>&1
  tail -3 gpurun_out/sanitizer_${tool}_r02.log
done
echo "== sanitizer on k-NN + tick (memcheck)"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_tick_r02.log python -m pytest tests/test_gpu_tick.py tests/test_gpu_knn.py -x -q -k "not large and not full" > gpurun_out/sanitizer_memcheck_tick_r02.out 2>&1
tail -3 gpurun_out/sanitizer_memcheck_tick_r02.log; tail -2 gpurun_out/sanitizer_memcheck_tick_r02.out
