cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py tests/test_edm_parity_gpu.py tests/test_sdxl_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
python profiles/detail.py > gpurun_out/s2_detail_gn8.txt 2>&1
head -1 gpurun_out/s2_detail_gn8.txt; grep gn_apply gpurun_out/s2_detail_gn8.txt | head -8
awk 'NR>1{t[$6]+=$1} END{for(k in t) printf "%s %.3f\n", k, t[k]}' gpurun_out/s2_detail_gn8.txt
