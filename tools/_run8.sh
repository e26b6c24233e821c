set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r2_pytest_8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_8.log
tail -6 gpurun_out/r2_pytest_8.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2_bench_8.log 2>&1; tail -c 2800 gpurun_out/r2_bench_8.log | head -c 1200
LASR_FUSED_FFN=0 timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2_bench_8b.log 2>&1; tail -c 2800 gpurun_out/r2_bench_8b.log | head -c 600
timeout 600 python tools/step_profile.py c2 0 45 > gpurun_out/r2_step_profile_1.log 2>&1
