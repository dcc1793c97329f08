#!/bin/bash
# post-processing change loop: bit-exact tests, then a short bench (headline + kernel_ms_per_step)
mkdir -p gpurun_out
bash scripts/gpu_ci.sh tests/test_postproc_gpu.py tests/test_pipeline_gpu.py
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-other-precision --steady-seconds 0 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench exit=$?"
tail -3 gpurun_out/bench_short.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_short.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step','postproc_parity') if k in d})
print('e2e',d['e2e']['value'], 'pw',d['roofline_pointwise']['frac'],'dw',d['roofline_depthwise']['frac'],'all',d['roofline']['frac'])
print({k:v for k,v in d['layer_ms'].items() if 'model.1' in k or 'model.2' in k or k=='model.0'})
PY
