#!/bin/bash
# Wan 1.3B self-forcing under sequence parallelism: NCCL all-to-all vs the one-pass peer exchange vs peer + CUDA graphs.
#   gpurun --gpus N --timeout 1200 -- 'bash profiles/wan_sp_modes.sh N'
N=${1:-2}
mkdir -p gpurun_out
port=29520
for mode in nccl peer peer_graphs; do
  export FASTGEN_B200_SP_EXCHANGE=${mode%%_*}
  if [ "$mode" = "peer_graphs" ]; then export FASTGEN_B200_SP_GRAPHS=1; else unset FASTGEN_B200_SP_GRAPHS; fi
  port=$((port + 1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --workload wan --gpus $N --steps 3 --warmup 2 > gpurun_out/r2_wan_n${N}_$mode.json 2> gpurun_out/r2_wan_n${N}_$mode.err
  echo "wan N=$N $mode rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_wan_n${N}_$mode.json").read().strip().splitlines()[-1])
    print("   ms_per_step", round(d["ms_per_step"], 1), "value", round(d["value"], 3), "launches", d["gpu_launches"])
except Exception as e:
    print("   no line:", e)
PY
done
