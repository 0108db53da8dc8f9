cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py tests/test_edm_parity_gpu.py tests/test_fullsize_parity_gpu.py tests/test_x3_f32_attention_gpu.py -m gpu -x -q -k "not wan and not Wan" 2>&1 | tail -2
python profiles/detail.py > gpurun_out/s2_detail_xds.txt 2>&1
head -1 gpurun_out/s2_detail_xds.txt; grep -E "res=3|res=1 |rs=3" gpurun_out/s2_detail_xds.txt | head -12
awk 'NR>1{t[$6]+=$1} END{for(k in t) printf "%s %.3f\n", k, t[k]}' gpurun_out/s2_detail_xds.txt
