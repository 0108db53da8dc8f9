cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py tests/test_transformer_kernels_gpu.py tests/test_wan_parity_gpu.py tests/test_dit_parity_gpu.py -m gpu -x -q 2>&1 | tail -2

python bench.py --workload wan --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s2_wan_lean2.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/s2_wan_lean2.json',):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    r=d['roofline']
    print(f, round(d['value'],4), d['ms_per_step'], r.get('gemm_tflops'), {k:v for k,v in r['per_kernel_ms'].items() if 'conv_gemm' in k})
PY
