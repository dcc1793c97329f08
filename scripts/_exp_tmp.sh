mkdir -p gpurun_out; : > gpurun_out/layers.jsonl
timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 -k "heads or network" > gpurun_out/test_net_gpu.log 2>&1; echo "net tests exit=$?"; tail -3 gpurun_out/test_net_gpu.log
for v in $VARIANTS; do
  LWP_ALLOW_TIMING_EXPERIMENTS=1 LWP_DEBUG_HEADS=$v timeout 300 python scripts/time_layers.py heads >> gpurun_out/layers.jsonl 2>gpurun_out/layers.err || { echo "variant $v failed"; tail -5 gpurun_out/layers.err; }
done
cat gpurun_out/layers.jsonl
