#!/bin/bash
# A/B of the round-1j density-kernel changes (lane-pair gather, 2 CTAs/SM MLP backward).  Run under gpurun.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j_pytest.log
python bench.py --no-cpu-baseline > $O/j_c2_new.json 2> $O/j_c2_new.err
NRC_QUERY_PAIR=0 python bench.py --no-cpu-baseline > $O/j_c2_pair0.json 2>/dev/null
NRC_MLP_BWD_GRID_MULT=1 python bench.py --no-cpu-baseline > $O/j_c2_bwd1.json 2>/dev/null
python bench.py --workload config3 --no-cpu-baseline > $O/j_c3_new.json 2> $O/j_c3_new.err
NRC_QUERY_PAIR=0 python bench.py --workload config3 --no-cpu-baseline > $O/j_c3_pair0.json 2>/dev/null
NRC_QUERY_PAIR=0 NRC_QUERY_GRID_MULT=3 python bench.py --workload config3 --no-cpu-baseline > $O/j_c3_pair0_m3.json 2>/dev/null
python tests/tools/bench_query.py --only mma > $O/j_query_pair1.log 2>&1
NRC_QUERY_PAIR=0 python tests/tools/bench_query.py --only mma > $O/j_query_pair0.log 2>&1
M=l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_read.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:mlp_bf16_fwd -c 8 --csv --log-file $O/j_ncu_l1_pair1.csv python tests/tools/bench_query.py --only mma --reps 1 > $O/j_ncu1.log 2>&1
NRC_QUERY_PAIR=0 ncu --metrics $M --clock-control none -k regex:mlp_bf16_fwd -c 8 --csv --log-file $O/j_ncu_l1_pair0.csv python tests/tools/bench_query.py --only mma --reps 1 > $O/j_ncu0.log 2>&1
tail -3 $O/j_pytest.log
for f in $O/j_c2_new.json $O/j_c2_pair0.json $O/j_c2_bwd1.json $O/j_c3_new.json $O/j_c3_pair0.json $O/j_c3_pair0_m3.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], {k:round(v,4) for k,v in d.get("kernel_ms",{}).items() if "query" in k or "mlp_bwd" in k})
except Exception as e: print("ERR", e)
PY
done
cat $O/j_query_pair1.log $O/j_query_pair0.log
