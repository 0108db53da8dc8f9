cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/s2_gpu_tests_full.log 2>&1; echo rc=$? >> gpurun_out/s2_gpu_tests_full.log
tail -4 gpurun_out/s2_gpu_tests_full.log
python bench.py --steps 10 --warmup 3 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['per_kernel_ms_per_step'], d['roofline'].get('model_flops_utilisation'))
print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d.get('also',{}).items()})
PY
