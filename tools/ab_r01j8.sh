#!/bin/bash
O=gpurun_out; N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/test_allreduce.py > $O/j8_allreduce_n$N.json 2> $O/j8_allreduce_n$N.err
NRC_ALLREDUCE=nccl timeout 300 $TR --master-port 29512 bench.py --gpus $N > $O/j8_bench_n${N}_nccl.json 2> $O/j8_bench_n${N}_nccl.err
timeout 300 $TR --master-port 29513 bench.py --gpus $N > $O/j8_bench_n${N}_peer.json 2> $O/j8_bench_n${N}_peer.err
tail -n 1 $O/j8_allreduce_n$N.json; tail -n 5 $O/j8_allreduce_n$N.err
for f in $O/j8_bench_n${N}_nccl.json $O/j8_bench_n${N}_peer.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["config"].get("allreduce"))
except Exception as e: print("ERR", e)
PY
done
tail -n 3 $O/j8_bench_n${N}_peer.err
