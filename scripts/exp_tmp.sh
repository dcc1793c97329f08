timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2
for v in 0 15; do LWP_DEBUG_GEMM=$v python scripts/time_layers.py model.1.pw model.2.pw model.3.pw model.4.pw cpm.align cpm.trunk.0.pw heads.1 refinement_stages.0.trunk.1.initial 2>&1 | tail -1 | cut -c1-500; done
