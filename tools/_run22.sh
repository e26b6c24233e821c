timeout 300 python tools/wgrad_probe.py 37674 256 2048 2>&1 | tail -8
timeout 300 python tools/wgrad_probe.py 37674 2048 256 2>&1 | tail -8
timeout 300 python tools/wgrad_probe.py 37674 256 256 2>&1 | tail -8
