timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -3
for v in "LWP_X=0" "LWP_DW_TMA_STORE=0"; do env $v python scripts/time_layers.py .dw 2>&1 | tail -1 | cut -c1-600; done
