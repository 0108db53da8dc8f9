cd $GRAFT_REPO_ROOT
for d in 0 256 272 368; do
BISECT=1 FG_CONV_DBG=$d python profiles/experiments/conv_epilogue_timing.py >> gpurun_out/s2_conv_epi_bisect2.txt 2>&1
done
cat gpurun_out/s2_conv_epi_bisect2.txt
python -m pytest tests/test_kernels_gpu.py tests/test_edm_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
