timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2
for v in "LWP_X=0"; do env $v python scripts/time_layers.py cpm.trunk.1.pw cpm.trunk.2.pw refinement_stages.0.trunk.1 2>/dev/null | cut -c1-800; done
