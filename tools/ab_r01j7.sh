#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/j7_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j7_pytest.log
python bench.py --no-cpu-baseline > $O/j7_c2.json 2>$O/j7_c2.err
NRC_QUERY_PPW=32 python bench.py --no-cpu-baseline > $O/j7_c2_ppw32.json 2>$O/j7_c2_ppw32.err
python bench.py --workload config3 --no-cpu-baseline > $O/j7_c3.json 2>$O/j7_c3.err
tail -n 3 $O/j7_pytest.log
for f in $O/j7_c*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d.get("gpu_launches_per_step"), {k:round(v,4) for k,v in sorted(d.get("kernel_ms",{}).items(), key=lambda kv:-kv[1])[:9]})
except Exception as e: print("ERR", e)
PY
done
