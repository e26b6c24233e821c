set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_pair_gpu.py -x -q --timeout=300 > gpurun_out/r2f_pair.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2f_pair.log
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest.log; tail -5 gpurun_out/r2f_pytest.log
(time timeout 600 python bench.py --steps 20 --warmup 5 --quick) > gpurun_out/r2f_bench_quick.log 2>&1; grep -E "^real|Traceback" gpurun_out/r2f_bench_quick.log
LASR_FUSED_ATTN_BWD=0 timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2f_bench_quick_nopair.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r2f_bench_quick.log', 'gpurun_out/r2f_bench_quick_nopair.log'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f, d['value'], d['ms_per_step'], d.get('e2e', {}).get('ms_per_step'), d.get('roofline', {}).get('frac'), d.get('extra'))
PY
