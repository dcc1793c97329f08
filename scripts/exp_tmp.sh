timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -5
for v in "LWP_X=0" "LWP_CONV3=0"; do env $v python scripts/time_layers.py initial_stage.trunk cpm.conv refinement_stages.0.trunk.1 2>&1 | tail -1 | cut -c1-700; done
