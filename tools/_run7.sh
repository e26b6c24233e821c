mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ffn_fused_gpu.py -q --timeout=300 -x 2>&1 | tail -12
timeout 120 python tools/ffn_bench.py 37674 256 2048 1 2>&1 | tail -1
timeout 120 python tools/ffn_bench.py 37674 256 2048 0 2>&1 | tail -1
timeout 120 python tools/ffn_trace.py 2>&1 | sed -n 8,22p
