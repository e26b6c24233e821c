set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/r2b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest.log; tail -5 gpurun_out/r2b_pytest.log
timeout 300 python tools/step_profile.py c2 0 70 0.1 > gpurun_out/r2b_step_profile_p1.log 2>&1
(time timeout 1200 python bench.py --steps 20 --warmup 5) > gpurun_out/r2b_bench.log 2>&1; grep -E "^real|Traceback" gpurun_out/r2b_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2b_step_traffic.csv python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
