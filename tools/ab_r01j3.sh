#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/j3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j3_pytest.log
for dual in 1 0; do
  NRC_QUERY_DUAL=$dual python bench.py --no-cpu-baseline > $O/j3_c2_dual$dual.json 2>$O/j3_c2_dual$dual.err
  NRC_QUERY_DUAL=$dual python bench.py --workload config3 --no-cpu-baseline > $O/j3_c3_dual$dual.json 2>$O/j3_c3_dual$dual.err
done
M=l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_read.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:mlp_bf16_fwd -c 8 --csv --log-file $O/j3_ncu_l1.csv python tests/tools/bench_query.py --only mma --reps 1 > $O/j3_ncu.log 2>&1
tail -n 3 $O/j3_pytest.log
for f in $O/j3_c*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], {k:round(v,4) for k,v in d.get("kernel_ms",{}).items() if "query" in k or "mlp_bwd" in k})
    r=d["roofline"]; print(r.get("l2_gather"), r.get("l2_gather_frac"), {k:v for k,v in r.items() if isinstance(v,dict) and "l2_gather_frac" in v})
except Exception as e: print("ERR", e)
PY
done
