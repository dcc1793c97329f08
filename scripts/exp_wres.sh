#!/bin/bash
# weight-resident thin 1x1 kernel: parity tests, then layer timings with the kernel off / default / forced
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 -k "wres" > gpurun_out/test_wres.log 2>&1; echo "wres tests exit=$?"; tail -15 gpurun_out/test_wres.log
: > gpurun_out/layers.jsonl
for v in "LWP_GEMM_WRES=0" "X=0" "LWP_GEMM_WRES=1" "LWP_GEMM_WRES=1 LWP_GEMM_STAGES=3"; do
  env $v timeout 300 python scripts/time_layers.py .pw cpm.align initial >> gpurun_out/layers.jsonl 2>gpurun_out/layers.err || { echo "variant $v failed"; tail -5 gpurun_out/layers.err; }
done
cat gpurun_out/layers.jsonl
