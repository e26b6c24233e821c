mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_pytest.log; tail -4 gpurun_out/r2n_pytest.log
b() { env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --quick $EXTRA 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$* $EXTRA', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['frac'],4), d.get('gpu_launches_per_step'))
"; }
b A=1
EXTRA="--batch 32" b A=b32
EXTRA="--batch 64" b A=b64
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2n_step.csv python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r2n_step.csv | grep -E "layernorm|total"
