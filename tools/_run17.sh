set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest.log; tail -8 gpurun_out/r2h_pytest.log
timeout 120 python tools/ffn_bench.py 37674 256 2048 1 fwd0 2>&1 | tail -1
timeout 120 python tools/ffn_bench.py 37674 256 2048 1 fwd0.1 2>&1 | tail -1
(time timeout 600 python bench.py --steps 20 --warmup 5 --quick) > gpurun_out/r2h_bench_quick.log 2>&1; grep -E "^real|Traceback" gpurun_out/r2h_bench_quick.log
python - <<'PY'
import json
for f in ('gpurun_out/r2h_bench_quick.log',):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f, d['value'], d['ms_per_step'], d.get('e2e', {}).get('ms_per_step'), d.get('roofline', {}).get('frac'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2h_step_traffic.csv python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"
