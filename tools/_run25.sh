mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_attn_fused_gpu.py tests/test_u2_gpu.py -q --timeout=600 2>&1 | tail -3
b() { env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['frac'],4))
"; }
b LASR_GEMM_ROT=1
b LASR_GEMM_ROT=0
b LASR_GEMM_ROT=1
b LASR_GEMM_ROT=0
