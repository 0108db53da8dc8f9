set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_kernels_gpu.py tests/test_edm_parity_gpu.py -m gpu -x -q > gpurun_out/s2_tests1.log 2>&1; echo rc=$? >> gpurun_out/s2_tests1.log
FG_CONV_SLABS=4 python profiles/detail.py > gpurun_out/s2_detail_slabs4.txt 2>&1
FG_CONV_SLABS=8 python profiles/detail.py > gpurun_out/s2_detail_slabs8.txt 2>&1
FG_CONV_SLABS=12 python profiles/detail.py > gpurun_out/s2_detail_slabs12.txt 2>&1
FG_CONV_SLABS=8 FG_CONV_SLABS_K=27 python profiles/detail.py > gpurun_out/s2_detail_slabs8_k27.txt 2>&1
timeout 120 python profiles/experiments/attn_phase_timing.py > gpurun_out/s2_attn_phase.txt 2>&1
grep -h "res=1" gpurun_out/s2_detail_slabs*.txt | sort -k9,12 | head -80
head -1 gpurun_out/s2_detail_slabs*.txt
tail -3 gpurun_out/s2_tests1.log
