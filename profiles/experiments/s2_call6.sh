cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py tests/test_edm_parity_gpu.py tests/test_transformer_kernels_gpu.py tests/test_sdxl_parity_gpu.py -m gpu -x -q > gpurun_out/s2_tests6.log 2>&1; echo rc=$? >> gpurun_out/s2_tests6.log
tail -5 gpurun_out/s2_tests6.log
python profiles/experiments/conv_epilogue_timing.py > gpurun_out/s2_conv_epi_timing_smemstats.txt 2>&1
cat gpurun_out/s2_conv_epi_timing_smemstats.txt
FG_CONV_SMEM_STATS=0 python profiles/detail.py > gpurun_out/s2_detail_c_smem0.txt 2>&1
FG_CONV_SMEM_STATS=1 python profiles/detail.py > gpurun_out/s2_detail_c_smem1.txt 2>&1
head -1 gpurun_out/s2_detail_c_smem*.txt
