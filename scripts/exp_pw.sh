#!/bin/bash
# timing-experiment decomposition of the thin 1x1 layers (needs a -DLWP_TIMING_EXPERIMENTS build)
mkdir -p gpurun_out
: > gpurun_out/layers.jsonl
export LWP_ALLOW_TIMING_EXPERIMENTS=1
for v in "X=0" "LWP_DEBUG_GEMM=1" "LWP_DEBUG_GEMM=8" "LWP_DEBUG_GEMM=16" "LWP_DEBUG_GEMM=32" "LWP_DEBUG_GEMM=48" "LWP_DEBUG_GEMM=9" "LWP_DEBUG_GEMM=15" "LWP_DEBUG_GEMM=5" "LWP_STAGING=2" "LWP_GEMM_STAGES=4" "LWP_ACC_STAGES=2" "LWP_ACC_STAGES=4"; do
  env $v timeout 300 python scripts/time_layers.py model.1.pw model.2.pw model.4.pw cpm.trunk cpm.align >> gpurun_out/layers.jsonl 2>gpurun_out/layers.err || { echo "variant $v failed"; tail -5 gpurun_out/layers.err; }
done
cat gpurun_out/layers.jsonl
