set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest.log; tail -5 gpurun_out/r2d_pytest.log
(time timeout 600 python bench.py --steps 20 --warmup 5 --quick) > gpurun_out/r2d_bench_quick.log 2>&1; grep -E "^real|Traceback" gpurun_out/r2d_bench_quick.log
python - <<'PY'
import json
for l in open('gpurun_out/r2d_bench_quick.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d.get('e2e'), d.get('roofline', {}).get('frac'))
PY
