cd $GRAFT_REPO_ROOT
for bx in 0 -1 8 4; do
echo "== FG_ATTN_BX=$bx" >> gpurun_out/s2_attn_bx.txt
FG_ATTN_BX=$bx python profiles/attn_micro.py 2>&1 | grep -E "edm|Error|error" >> gpurun_out/s2_attn_bx.txt
FG_ATTN_BX=$bx python -m pytest tests/test_kernels_gpu.py tests/test_transformer_kernels_gpu.py -m gpu -x -q -k "attn or attention" 2>&1 | tail -2 >> gpurun_out/s2_attn_bx.txt
done
cat gpurun_out/s2_attn_bx.txt
