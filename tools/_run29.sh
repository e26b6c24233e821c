timeout 300 python -m pytest tests/test_attn_pair_gpu.py -q --timeout=300 2>&1 | tail -2
timeout 200 python tools/dv_probe.py 2>&1 | tail -4
