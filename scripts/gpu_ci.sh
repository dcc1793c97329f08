#!/bin/bash
# Runs on the GPU box (via gpurun): every GPU test file in its own process so that one CUDA fault
# cannot poison the others; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  b=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q -x --timeout 600 > "gpurun_out/$b.log" 2>&1
  echo "$b exit=$?" | tee -a gpurun_out/summary.txt
  tail -5 "gpurun_out/$b.log"
done
