#!/bin/bash
O=gpurun_out
T=${1:-j6}
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
python bench.py --no-cpu-baseline > $O/${T}_c2.json 2>$O/${T}_c2.err
python bench.py --workload config3 --no-cpu-baseline > $O/${T}_c3.json 2>$O/${T}_c3.err
tail -n 3 $O/${T}_pytest.log
for f in $O/${T}_c*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], {k:round(v,4) for k,v in sorted(d.get("kernel_ms",{}).items(), key=lambda kv:-kv[1])[:8]})
except Exception as e: print("ERR", e)
PY
done
