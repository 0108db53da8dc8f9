cd $GRAFT_REPO_ROOT
for d in 0 16 48 112 128 240; do
BISECT=1 FG_CONV_DBG=$d python profiles/experiments/conv_epilogue_timing.py >> gpurun_out/s2_conv_epi_bisect.txt 2>&1
done
cat gpurun_out/s2_conv_epi_bisect.txt
