#!/bin/bash
O=gpurun_out; N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29513 bench.py --gpus $N > $O/j10_bench_n${N}_ingraph.json 2> $O/j10_bench_n${N}_ingraph.err
NRC_AR_CTAS_OVERLAP=32 timeout 300 $TR --master-port 29514 bench.py --gpus $N > $O/j10_bench_n${N}_ingraph32.json 2> $O/j10_bench_n${N}_ingraph32.err
NRC_AR_CTAS_OVERLAP=296 timeout 300 $TR --master-port 29515 bench.py --gpus $N > $O/j10_bench_n${N}_ingraph296.json 2> $O/j10_bench_n${N}_ingraph296.err
NRC_DP_INGRAPH=0 timeout 300 $TR --master-port 29516 bench.py --gpus $N > $O/j10_bench_n${N}_after.json 2> $O/j10_bench_n${N}_after.err
for f in $O/j10_bench_n${N}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["config"].get("allreduce"))
except Exception as e: print("ERR", e)
PY
done
tail -n 6 $O/j10_bench_n${N}_ingraph.err
