cd $GRAFT_REPO_ROOT
python -m pytest tests/test_wan_sp_gpu.py tests/test_unverified_paths_gpu.py -m gpu -x -q > gpurun_out/s2_gpu_tests_n2.log 2>&1; echo rc=$? >> gpurun_out/s2_gpu_tests_n2.log
tail -4 gpurun_out/s2_gpu_tests_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s2_bench_n2.json 2> gpurun_out/s2_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
print({k:(round(v.get('value'),4) if isinstance(v,dict) and v.get('value') else None) for k,v in d.get('also',{}).items()})
PY
