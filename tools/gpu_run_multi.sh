#!/bin/bash
# multi-GPU: the contract bench under torchrun (weak scaling + the one-stream NCCL scatter/gather leg), levels 10 and 41
N=${1:-2}
TAG=${2:-r02m}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi topo -m 2>/dev/null | head -12
for lvl in 10 41; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --level $lvl --legs "" > gpurun_out/${TAG}_n${N}_l${lvl}.json 2> gpurun_out/${TAG}_n${N}_l${lvl}.err
  tail -5 gpurun_out/${TAG}_n${N}_l${lvl}.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_n${N}_l${lvl}.json").read().strip().split("\n")[-1])
    print("N=$N level $lvl: value", d["value"], "e2e", d.get("e2e",{}).get("value"), "frac_of_ceiling", d.get("e2e",{}).get("fraction_of_bare_copy_ceiling"))
    o=d.get("one_stream")
    if o: print("  one_stream:", o["value"], "MB/s ok", o["round_trip_ok"], "ms", o["ms_per_step"], o["phases_ms"], o["leg_GBps_rank0_link"])
except Exception as ex: print("parse failed", ex)
PY
  el level$lvl
done
