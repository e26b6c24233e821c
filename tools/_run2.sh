set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r2_pytest_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_2.log
tail -40 gpurun_out/r2_pytest_2.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2a.log 2>&1; tail -c 1500 gpurun_out/r2_bench_2a.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dropout 0.1 > gpurun_out/r2_bench_2b.log 2>&1; tail -c 1500 gpurun_out/r2_bench_2b.log
python tools/gemm_shapes.py > gpurun_out/r2_gemm_shapes_1.log 2>&1; tail -5 gpurun_out/r2_gemm_shapes_1.log
