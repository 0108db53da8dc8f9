cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py tests/test_edm_parity_gpu.py tests/test_transformer_kernels_gpu.py tests/test_sdxl_parity_gpu.py tests/test_dit_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
FG_CONV_PAIR=0 python profiles/detail.py > gpurun_out/s2_detail_pair0.txt 2>&1
FG_CONV_PAIR=1 python profiles/detail.py > gpurun_out/s2_detail_pair1.txt 2>&1
head -1 gpurun_out/s2_detail_pair*.txt
for f in gpurun_out/s2_detail_pair0.txt gpurun_out/s2_detail_pair1.txt; do awk 'NR>1{t[$6]+=$1} END{for(k in t) printf "%s %.3f\n", k, t[k]}' $f | grep conv; grep -E "K=2112 N=192|K=384 N=1152|K=1728 N=192 out=bfloat16|K=72 N=192|K=4224" $f | head -5; done
