#!/bin/bash
O=gpurun_out; N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29511 tools/test_allreduce.py > $O/j11_allreduce_n$N.json 2> $O/j11_allreduce_n$N.err
timeout 240 $TR --master-port 29513 bench.py --gpus $N > $O/j11_bench_n${N}.json 2> $O/j11_bench_n${N}.err
tail -n 1 $O/j11_allreduce_n$N.json; tail -n 4 $O/j11_allreduce_n$N.err
python - "$O/j11_bench_n${N}.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["config"].get("allreduce"))
except Exception as e: print("ERR", e)
PY
tail -n 4 $O/j11_bench_n${N}.err
