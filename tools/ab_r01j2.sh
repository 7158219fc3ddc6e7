#!/bin/bash
O=gpurun_out
for occ in 4 5 6; do
  NRC_QUERY_OCC=$occ python bench.py --no-cpu-baseline > $O/j2_c2_occ$occ.json 2>/dev/null
  NRC_QUERY_OCC=$occ python bench.py --workload config3 --no-cpu-baseline > $O/j2_c3_occ$occ.json 2>/dev/null
  NRC_QUERY_OCC=$occ python tests/tools/bench_query.py --only mma > $O/j2_query_occ$occ.log 2>&1
done
for f in $O/j2_c*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], {k:round(v,4) for k,v in d.get("kernel_ms",{}).items() if "query" in k or "mlp_bwd" in k})
except Exception as e: print("ERR", e)
PY
done
tail -n 3 $O/j2_query_occ*.log
