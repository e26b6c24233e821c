mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_planes_gpu.py tests/test_wgrad2_gpu.py -x -q --timeout=300 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_pytest.log; tail -3 gpurun_out/r2v_pytest.log
b() { env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['frac'],4))
"; }
b LASR_CONV2_WGRAD2=1
b LASR_CONV2_WGRAD2=0
b LASR_CONV2_WGRAD2=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2v_step.csv python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2v_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2v_step.csv | head -30
