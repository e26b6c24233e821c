set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_pytest.log; tail -8 gpurun_out/r2j_pytest.log
b() { env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['frac'],4))
"; }
b A=1
b LASR_GEMM_FILL=0
b LASR_FUSED_ATTN_BWD=0
b A=2
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2j_step_traffic.csv python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2j_ncu.log 2>&1; echo "ncu rc=$?"
