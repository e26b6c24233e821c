mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_wgrad2_gpu.py -x -q --timeout=300 2>&1 | tail -4
timeout 200 python tools/wgrad2_probe.py 37674 256 2048 2>&1 | tail -4
timeout 200 python tools/wgrad2_probe.py 37674 256 256 2>&1 | tail -4
timeout 200 python tools/wgrad2_probe.py 37674 768 256 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_pytest.log; tail -4 gpurun_out/r2s_pytest.log
b() { env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['frac'],4))
"; }
b LASR_WGRAD2=1
b LASR_WGRAD2=0
b LASR_WGRAD2=1
b LASR_WGRAD2=0
