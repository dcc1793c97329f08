#!/bin/bash
# tests -> smoke -> bench -> ncu launch list (each under its own timeout)
mkdir -p gpurun_out
bash scripts/gpu_ci.sh tests/test_postproc_gpu.py tests/test_net_gpu.py tests/test_pipeline_gpu.py
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
