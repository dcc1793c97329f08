#!/bin/bash
mkdir -p gpurun_out
for d in 2 3 4; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-other-precision --steady-seconds 0 --no-roofline --e2e-depth $d > gpurun_out/b.json 2>gpurun_out/b.err || { echo "depth $d failed"; tail -5 gpurun_out/b.err; }
  python - "$d" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1])
print('depth',sys.argv[1], 'value',round(d['value']), 'e2e',round(d['e2e']['value']), 'f32',round(d['e2e_f32']['value']), 'raw',round(d['e2e_raw_frames']['value']))
PY
done
