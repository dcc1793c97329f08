#!/bin/bash
mkdir -p gpurun_out
for v in "LWP_PAF_BLOCKS=1" "LWP_PAF_BLOCKS=2" "LWP_PAF_BLOCKS=3" "LWP_PAF_BLOCKS=4" "LWP_PAF_BLOCKS=8" "LWP_NO_PAF_PACK=1"; do
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-other-precision --steady-seconds 0 --no-u8 > gpurun_out/b.json 2>gpurun_out/b.err || { echo "variant $v failed"; tail -5 gpurun_out/b.err; }
  python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'])
PY
done
