set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ffn_fused_gpu.py -q --timeout=300 -x > gpurun_out/r2_pytest_6a.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_6a.log; tail -15 gpurun_out/r2_pytest_6a.log
timeout 300 python tools/ffn_bench.py > gpurun_out/r2_ffn_bench_3.log 2>&1; cat gpurun_out/r2_ffn_bench_3.log
