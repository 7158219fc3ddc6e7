#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/j5_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j5_pytest.log
for agg in 1 0; do
  NRC_ENC_BWD_AGG=$agg python bench.py --no-cpu-baseline > $O/j5_c2_agg$agg.json 2>$O/j5_c2_agg$agg.err
done
python bench.py --workload config3 --no-cpu-baseline > $O/j5_c3.json 2>$O/j5_c3.err
tail -n 3 $O/j5_pytest.log
for f in $O/j5_c*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], {k:round(v,4) for k,v in d.get("kernel_ms",{}).items() if "query" in k or "mlp_bwd" in k or "encode" in k})
except Exception as e: print("ERR", e)
PY
done
