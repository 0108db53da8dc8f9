cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s2_bench_n2c.json 2> gpurun_out/s2_bench_n2c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_n2c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
print({k:(round(v.get('value'),4) if isinstance(v,dict) and v.get('value') else None) for k,v in d.get('also',{}).items()})
PY
