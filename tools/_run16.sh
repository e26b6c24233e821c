set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ffn_fused_gpu.py -x -q --timeout=300 > gpurun_out/r2g_ffn.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2g_ffn.log
timeout 120 python tools/ffn_bench.py 37674 256 2048 1 fwd0 2>&1 | tail -1
timeout 120 python tools/ffn_bench.py 37674 256 2048 1 fwd0.1 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest.log; tail -5 gpurun_out/r2g_pytest.log
(time timeout 600 python bench.py --steps 20 --warmup 5 --quick) > gpurun_out/r2g_bench_quick.log 2>&1; grep -E "^real|Traceback" gpurun_out/r2g_bench_quick.log
LASR_FUSED_FFN_FWD=0 timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2g_bench_quick_nofwd.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r2g_bench_quick.log', 'gpurun_out/r2g_bench_quick_nofwd.log'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f, d['value'], d['ms_per_step'], d.get('e2e', {}).get('ms_per_step'), d.get('roofline', {}).get('frac'))
PY
