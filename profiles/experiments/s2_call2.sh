set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py tests/test_edm_parity_gpu.py tests/test_fullsize_parity_gpu.py -m gpu -x -q -k "not wan and not Wan" > gpurun_out/s2_tests2.log 2>&1; echo rc=$? >> gpurun_out/s2_tests2.log
python profiles/experiments/conv_epilogue_timing.py > gpurun_out/s2_conv_epi_timing.txt 2>&1
FASTGEN_B200_BF16_H1=0 python profiles/detail.py > gpurun_out/s2_detail_h1f32.txt 2>&1
FASTGEN_B200_BF16_H1=1 python profiles/detail.py > gpurun_out/s2_detail_h1bf16.txt 2>&1
head -1 gpurun_out/s2_detail_h1*.txt
tail -3 gpurun_out/s2_tests2.log
cat gpurun_out/s2_conv_epi_timing.txt
