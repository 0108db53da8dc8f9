cd $GRAFT_REPO_ROOT
run() { echo "== $*" >> gpurun_out/s2_attn_tp.txt; env "$@" timeout 300 python profiles/attn_micro.py 2>&1 | grep -E "edm|chunk6 self  tq|bidir|SP4|Error|error" >> gpurun_out/s2_attn_tp.txt; }
run FG_ATTN_TP=0
run FG_ATTN_TP=1
run FG_ATTN_TP=1 FG_ATTN_POLY=0
run FG_ATTN_TP=1 FG_ATTN_POLY=4
FG_ATTN_TP=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_transformer_kernels_gpu.py -m gpu -x -q -k "attn or attention" 2>&1 | tail -2 >> gpurun_out/s2_attn_tp.txt
cat gpurun_out/s2_attn_tp.txt
