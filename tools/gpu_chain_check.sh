#!/bin/bash
# chain kernel v2 (CTA pairs, resident weights) against v1: parity tests, then device time per stack
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_gpu.py -x -q > gpurun_out/chain_v2_pytest.log 2>&1; echo "pytest v2 rc=$?" 
tail -5 gpurun_out/chain_v2_pytest.log
for P in 32768 524288; do
  timeout 300 python tools/bench_chain.py $P > gpurun_out/chain_v2_bench_$P.log 2>&1; echo "bench v2 $P rc=$?"; cat gpurun_out/chain_v2_bench_$P.log | grep -v Warn
  NRC_CHAIN_V1=1 timeout 300 python tools/bench_chain.py $P > gpurun_out/chain_v1_bench_$P.log 2>&1; echo "bench v1 $P rc=$?"; cat gpurun_out/chain_v1_bench_$P.log | grep -v Warn
done
