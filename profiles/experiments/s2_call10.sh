cd $GRAFT_REPO_ROOT
run() { echo "== $*" >> gpurun_out/s2_attn_pp.txt; env "$@" timeout 300 python profiles/attn_micro.py 2>&1 | grep -E "edm|chunk6 self  tq|bidir|Error|error" >> gpurun_out/s2_attn_pp.txt; }
run FG_ATTN_PP=0
run FG_ATTN_PP=1
run FG_ATTN_PP=1 FG_ATTN_POLY=0
run FG_ATTN_PP=1 FG_ATTN_POLY=4
run FG_ATTN_PP=1 FG_ATTN_BX=-1
FG_ATTN_PP=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_transformer_kernels_gpu.py -m gpu -x -q -k "attn or attention" 2>&1 | tail -2 >> gpurun_out/s2_attn_pp.txt
cat gpurun_out/s2_attn_pp.txt
