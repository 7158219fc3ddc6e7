#!/bin/bash
# per-op timelines of the chain kernel (trace build beside the product build)
export NRC_LIB_PATH=/root/repo/neural_radiance_caching_b200/libnrc_trace.so NRC_EXTRA_NVCC_FLAGS=-DNRC_CHAIN_TRACE
mkdir -p gpurun_out
for args in "$@"; do
  echo "=== $args"
  timeout 120 python tools/trace_chain.py $args 2>&1 | grep -v Warn | tee gpurun_out/trace_$(echo $args | tr ' ' '_').log
done
