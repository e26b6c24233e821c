set -x
mkdir -p gpurun_out
(time timeout 1500 python bench.py --steps 20 --warmup 5) > gpurun_out/r2t_bench.log 2>&1; grep -E "^real|Traceback" gpurun_out/r2t_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2t_step_traffic.csv python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2t_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2t_smoke.log
