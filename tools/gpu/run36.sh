for L in "" avoid-mpc_b200/lib/variants/libampc_lb6.so; do
  AMPC_LIB=$L timeout 600 python bench.py --mode knn_sweep --steps 5 2> gpurun_out/b36.err | python -c "
import json,sys
k=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('lib=[$L]', ' '.join('%d:%.3f/%.3f' % (r['npts'], r['search_ms'], r['stage_frac']) for r in k['rows']))"
done
