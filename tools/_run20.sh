set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_pytest.log; tail -6 gpurun_out/r2k_pytest.log
timeout 120 python tools/ffn_bench.py 37674 256 2048 1 fwd0 2>&1 | tail -1
timeout 120 python tools/ffn_bench.py 37674 256 2048 1 fwd0.1 2>&1 | tail -1
b() { env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['frac'],4))
"; }
b A=1
b A=2
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2k_step_traffic.csv python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2k_ncu.log 2>&1; echo "ncu rc=$?"
