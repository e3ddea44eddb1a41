for q in 2 8; do AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_ab.py --batch 96 2>&1 | tail -1; done
for b in 2048 4096 8192 16384 32768; do
  for k in warp quad; do
    for q in 4 8; do
      if [ $k = warp ] && [ $q = 4 ]; then continue; fi
      echo "kernel=$k Q=$q B=$b"; AMPC_SOLVE_KERNEL=$k AMPC_QUADS_PER_WARP=$q timeout 300 python tools/solve_only.py --streams 1 --steps 3 --batch $b --npts 4096 2>>gpurun_out/q9_err.log | tee gpurun_out/q9_${k}_q${q}_b${b}.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['stage_ms']['solve'], d['solves_per_s'])"
    done
  done
done
tail -3 gpurun_out/q9_err.log
