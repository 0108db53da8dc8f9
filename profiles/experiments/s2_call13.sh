cd $GRAFT_REPO_ROOT
python -m pytest tests/test_handoff_gpu.py tests/test_abi.py -m "gpu or not gpu" -x -q 2>&1 | tail -8
python profiles/membound_micro.py --handoff > gpurun_out/s2_membound_handoff.txt 2>&1
tail -8 gpurun_out/s2_membound_handoff.txt
