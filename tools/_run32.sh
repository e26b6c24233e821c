timeout 120 python tools/wgrad2_probe.py 1000 256 256 1 2>&1 | tail -6
timeout 200 python tools/wgrad2_probe.py 37674 256 2048 2>&1 | tail -9
