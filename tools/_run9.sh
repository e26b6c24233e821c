set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_u2_gpu.py tests/test_ctc_gpu.py -q --timeout=600 > gpurun_out/r2_pytest_9.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_9.log; tail -8 gpurun_out/r2_pytest_9.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_9.log 2>&1; grep -E "^real|Traceback" gpurun_out/r2_bench_9.log
