#!/bin/bash
# First GPU calls of the next round: everything written after round 1's GPU budget ended, in the order that isolates
# failures.  Each step has its own timeout; outputs land in gpurun_out/ (merged back by gpurun).
#
#   gpurun --timeout 1500 -- 'bash profiles/first_gpu_calls.sh 1'            # one GPU
#   gpurun --gpus 2 --timeout 1500 -- 'bash profiles/first_gpu_calls.sh 2'   # two GPUs: the peer exchange
mkdir -p gpurun_out
if [ "${1:-1}" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
  FASTGEN_B200_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_unverified_paths_gpu.py -m gpu -x -q \
    > gpurun_out/r2_unverified.log 2>&1; echo "unverified paths rc=$?"
  timeout 300 python profiles/membound_micro.py --vec > gpurun_out/r2_membound_vec.txt 2>&1; echo "membound rc=$?"
  timeout 300 python profiles/attn_micro.py --f32 > gpurun_out/r2_attn_micro_f32.txt 2>&1; echo "attn micro rc=$?"
else
  export FASTGEN_B200_TEST_UNVERIFIED=1
  timeout 300 python -m pytest tests/test_wan_sp_gpu.py -m gpu -x -q -k kernels_alone > gpurun_out/r2_peer_kernels.log 2>&1
  echo "peer kernels rc=$?"
  timeout 600 python -m pytest tests/test_wan_sp_gpu.py -m gpu -x -q -k "matches_single_gpu" > gpurun_out/r2_peer_wan.log 2>&1
  echo "peer wan rc=$?"
  timeout 600 python -m pytest tests/test_wan_sp_gpu.py -m gpu -x -q -k graphs > gpurun_out/r2_peer_graphs.log 2>&1
  echo "peer graphs rc=$?"
  for mode in nccl peer; do
    FASTGEN_B200_SP_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --workload wan --gpus 2 --steps 2 --warmup 1 \
      > gpurun_out/r2_bench_wan_n2_$mode.json 2> gpurun_out/r2_bench_wan_n2_$mode.err; echo "bench wan $mode rc=$?"
  done
  FASTGEN_B200_SP_EXCHANGE=peer FASTGEN_B200_SP_GRAPHS=1 timeout 600 python -m torch.distributed.run --nnodes=1 \
    --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload wan --gpus 2 --steps 2 --warmup 1 \
    > gpurun_out/r2_bench_wan_n2_peer_graphs.json 2> gpurun_out/r2_bench_wan_n2_peer_graphs.err; echo "bench wan peer+graphs rc=$?"
fi
