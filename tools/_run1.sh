set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r2_pytest_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_1.log
tail -5 gpurun_out/r2_pytest_1.log
python tools/ctc_bench.py > gpurun_out/r2_ctc_sweep_1.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1.log 2>&1; tail -c 3000 gpurun_out/r2_bench_1.log
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_step_traffic_1.csv python tools/one_step.py c2 > gpurun_out/r2_ncu_step_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 12 -c 2 -o gpurun_out/r2_fc1_swish python tools/gemm_bench.py --eager fc1 > gpurun_out/r2_ncu_fc1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 12 -c 2 -o gpurun_out/r2_dswish python tools/gemm_bench.py --eager dswish > gpurun_out/r2_ncu_dswish.log 2>&1
ls -la gpurun_out | tail -12
