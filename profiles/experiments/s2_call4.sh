cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py tests/test_edm_parity_gpu.py tests/test_transformer_kernels_gpu.py -m gpu -x -q > gpurun_out/s2_tests5.log 2>&1; echo rc=$? >> gpurun_out/s2_tests5.log
tail -5 gpurun_out/s2_tests5.log
python profiles/experiments/conv_epilogue_timing.py > gpurun_out/s2_conv_epi_timing_lean2.txt 2>&1
cat gpurun_out/s2_conv_epi_timing_lean2.txt
FG_CONV_LEAN=0 python profiles/detail.py > gpurun_out/s2_detail_b_lean0.txt 2>&1
FG_CONV_LEAN=1 python profiles/detail.py > gpurun_out/s2_detail_b_lean1.txt 2>&1
head -1 gpurun_out/s2_detail_b_lean*.txt
