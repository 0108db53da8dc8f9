cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/s2_gpu_tests_final.log 2>&1; echo rc=$? >> gpurun_out/s2_gpu_tests_final.log
tail -4 gpurun_out/s2_gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/s2_bench_final.json 2> gpurun_out/s2_bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup')}, d['e2e']['value'], round(d['roofline']['frac'],3), d['roofline']['per_kernel_ms_per_step'], round(d['roofline'].get('model_flops_utilisation'),3), d['clocks'])
print({k:(round(v.get('value'),4) if isinstance(v,dict) and v.get('value') else v.get('value') if isinstance(v,dict) else v) for k,v in d.get('also',{}).items()})
print(d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s2_bench_ref_final.json 2>/dev/null; tail -c 600 gpurun_out/s2_bench_ref_final.json
