import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f solves/s  ms/step %.3f  e2e %.0f  single-stream %s streams %s'%(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("single_stream",{}).get("value"), d["config"].get("streams")))
r=d["roofline"]; print('index: %.3f ms  %.0f GB/s  frac %.3f share %.2f'%(r['avg_launch_ms'],r['achieved'],r['frac'],r['step_share'])); print('knn stage:', {k:(round(v,4) if isinstance(v,float) else v) for k,v in r['knn_stage'].items() if k!='what'})
r=d["roofline_nlp"]; print('solve: %.3f ms  %.3f TF frac %.4f share %.2f'%(r['avg_launch_ms'],r['achieved'],r['frac'],r['step_share']))
print(d["latency"], d["solver"]); print(d.get("cpu_baseline")); print(d.get("clocks"), 'launches', d.get('gpu_launches'))
