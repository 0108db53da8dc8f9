cd $GRAFT_REPO_ROOT
FG_ATTN_SEP=1 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_transformer_kernels_gpu.py -m gpu -x -q -k "attn or attention" 2>&1 | tail -3 > gpurun_out/s2_attn_sep.txt
for e in "FG_ATTN_SEP=0" "FG_ATTN_SEP=1" "FG_ATTN_SEP=1 FG_ATTN_POLY=0" "FG_ATTN_SEP=1 FG_ATTN_POLY=4"; do echo "== $e" >> gpurun_out/s2_attn_sep.txt; env $e timeout 200 python profiles/attn_micro.py 2>&1 | grep -E "edm|Error|error" >> gpurun_out/s2_attn_sep.txt; done
cat gpurun_out/s2_attn_sep.txt
