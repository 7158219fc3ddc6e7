#!/bin/bash
# data-parallel A/B at N GPUs: bucket count x all-reduce kernel (usage under gpurun --gpus N: bash tools/scale_ab.sh N tag)
N=$1; T=${2:-r02}
mkdir -p gpurun_out
port=29600
for cfg in "3:auto" "2:auto" "3:peer-only" "3:multicast"; do
  b=${cfg%%:*}; m=${cfg#*:}
  port=$((port+1))
  env NRC_AR_BUCKETS=$b $( [ $m != auto ] && echo NRC_ALLREDUCE=$m ) timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-others > gpurun_out/${T}_n${N}_b${b}_${m}.json 2> gpurun_out/${T}_n${N}_b${b}_${m}.err
  echo "N=$N buckets=$b mode=$m rc=$?"
  python - <<PY
import json
txt=open("gpurun_out/${T}_n${N}_b${b}_${m}.json").read()
ls=[l for l in txt.splitlines() if l.startswith("{")]
if ls:
    d=json.loads(ls[-1]); print("  ms/step", round(d["ms_per_step"],4), "samples/s", round(d["value"]/1e6,1), "M |", d["config"]["allreduce"])
PY
done
