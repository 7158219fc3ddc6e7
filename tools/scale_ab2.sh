#!/bin/bash
# data-parallel A/B at N GPUs under environment switches.  usage (gpurun --gpus N): bash tools/scale_ab2.sh N TAG "ENV=.." "-" ...
N=$1; T=$2; shift 2
mkdir -p gpurun_out
port=29700; i=0
for v in "$@"; do
  i=$((i+1)); port=$((port+1)); e="$v"; [ "$v" = "-" ] && e=""
  env $e timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 20 --warmup 5 --no-others > gpurun_out/${T}_n${N}_v$i.json 2> gpurun_out/${T}_n${N}_v$i.err
  rc=$?
  python - "gpurun_out/${T}_n${N}_v$i.json" "$v" $rc <<'PY'
import json, sys
ls = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if ls:
    d = json.loads(ls[-1]); print(f"[{sys.argv[2]}] rc={sys.argv[3]} ms/step {d['ms_per_step']:.4f}  {d['value']/1e6:.1f} M samples/s")
else:
    print(f"[{sys.argv[2]}] rc={sys.argv[3]} no line")
PY
done
