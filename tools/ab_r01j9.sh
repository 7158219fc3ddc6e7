#!/bin/bash
O=gpurun_out; N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/test_allreduce.py > $O/j9_allreduce_n$N.json 2> $O/j9_allreduce_n$N.err
timeout 300 $TR --master-port 29513 bench.py --gpus $N > $O/j9_bench_n${N}_peer.json 2> $O/j9_bench_n${N}_peer.err
NRC_ALLREDUCE=multicast timeout 300 $TR --master-port 29514 bench.py --gpus $N > $O/j9_bench_n${N}_mc.json 2> $O/j9_bench_n${N}_mc.err
NRC_ALLREDUCE=nccl timeout 300 $TR --master-port 29515 bench.py --gpus $N > $O/j9_bench_n${N}_nccl1.json 2> $O/j9_bench_n${N}_nccl1.err
NRC_DP_OVERLAP=1 timeout 300 $TR --master-port 29516 bench.py --gpus $N > $O/j9_bench_n${N}_peer_overlap.json 2> $O/j9_bench_n${N}_peer_overlap.err
tail -n 1 $O/j9_allreduce_n$N.json
for f in $O/j9_bench_n${N}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["config"].get("allreduce"), d["config"]["parallelism"][:60])
except Exception as e: print("ERR", e)
PY
done
