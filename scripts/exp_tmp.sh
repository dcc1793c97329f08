timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2
for v in "LWP_X=0" "LWP_KBPS=1" "LWP_DEBUG_GEMM=15"; do env $v python scripts/time_layers.py initial_stage.trunk.1 cpm.align cpm.conv refinement_stages.0.trunk.1 model.3.pw model.2.pw heads 2>/dev/null | cut -c1-800; done
