set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 -k "dropout or optim or trainer_run" > gpurun_out/r2_pytest_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_3.log
tail -5 gpurun_out/r2_pytest_3.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_3.log 2>&1; tail -c 6000 gpurun_out/r2_bench_3.log
