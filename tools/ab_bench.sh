#!/bin/bash
# A/B of the config-2 line under environment switches.  usage (under gpurun): bash tools/ab_bench.sh TAG "ENV1=.. ENV2=.." "ENV=.." ...
# Each quoted argument is one variant (use "-" for the default environment); prints ms/step and the per-call kernel table.
T=$1; shift; O=gpurun_out; mkdir -p $O
i=0
for v in "$@"; do
  i=$((i+1)); e="$v"; [ "$v" = "-" ] && e=""
  env $e timeout 240 python bench.py --no-others --no-cpu-baseline > $O/${T}_v$i.json 2> $O/${T}_v$i.err; rc=$?
  python - "$O/${T}_v$i.json" "$v" $rc <<'PY'
import json, sys
try:
    l = [x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")][-1]
    l = json.loads(l)
    km = l.get("kernel_ms", {})
    top = sorted(km.items(), key=lambda kv: -kv[1])[:9]
    print(f"[{sys.argv[2]}] rc={sys.argv[3]} ms/step {l['ms_per_step']:.4f}  " + "  ".join(f"{k[4:]} {v*1e3:.1f}" for k, v in top))
except Exception as ex:
    print(f"[{sys.argv[2]}] rc={sys.argv[3]} failed: {ex}")
PY
done
