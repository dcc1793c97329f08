timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 -k "fused_heads" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -3
python scripts/time_layers.py heads 2>&1 | tail -1 | cut -c1-400
LWP_HEADS_FUSION=0 python scripts/time_layers.py heads 2>&1 | tail -1 | cut -c1-400
