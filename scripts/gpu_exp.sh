#!/bin/bash
# experiment loop on the GPU box: network parity tests, then layer timings under the given env variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/test_net_gpu.log 2>&1; echo "net tests exit=$?"; tail -3 gpurun_out/test_net_gpu.log
: > gpurun_out/layers.jsonl
for v in "$@"; do
  env $v timeout 300 python scripts/time_layers.py ${LAYERS:-dwpw} >> gpurun_out/layers.jsonl 2>gpurun_out/layers.err || { echo "variant $v failed"; tail -5 gpurun_out/layers.err; }
done
cat gpurun_out/layers.jsonl
