set -x
mkdir -p gpurun_out
for bo in 0 32 100 300; do
echo "backoff $bo"; LASR_GEMM_BACKOFF=$bo timeout 120 python tools/ffn_bench.py 37674 256 2048 1 fwd0.1 2>&1 | tail -1
done
b() { env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['frac'],4))
"; }
b LASR_GEMM_BACKOFF=0
b LASR_GEMM_BACKOFF=100
b LASR_GEMM_BACKOFF=0
b LASR_GEMM_BACKOFF=300
