set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ffn_fused_gpu.py -q --timeout=300 -x > gpurun_out/r2_pytest_4a.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_4a.log; tail -15 gpurun_out/r2_pytest_4a.log
timeout 300 python tools/ffn_bench.py > gpurun_out/r2_ffn_bench_1.log 2>&1; cat gpurun_out/r2_ffn_bench_1.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r2_pytest_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_4.log
tail -25 gpurun_out/r2_pytest_4.log
timeout 600 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2_bench_4.log 2>&1; tail -c 2500 gpurun_out/r2_bench_4.log
