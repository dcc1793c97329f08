timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2
for v in 0 8; do LWP_DEBUG_GEMM=$v python scripts/time_layers.py initial_stage.trunk cpm.align model.3.pw model.2.pw model.1.pw heads refinement_stages.0.trunk.1 2>/dev/null | cut -c1-900; done
