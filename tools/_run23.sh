timeout 300 python tools/small_gemm_probe.py 2>&1 | tail -16
LASR_PDL=0 timeout 300 python tools/small_gemm_probe.py 2>&1 | tail -16
