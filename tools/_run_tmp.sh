mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=800 > gpurun_out/r2_final_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_pytest.log; tail -3 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke gpurun_out/r2_final_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_final_step_traffic.csv python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2_final_ncu.log 2>&1; echo "ncu rc=$?"
python tools/traffic_summary.py gpurun_out/r2_final_step_traffic.csv --json gpurun_out/r2_traffic.json > gpurun_out/r2_step_traffic_final.txt; head -12 gpurun_out/r2_step_traffic_final.txt
cp gpurun_out/r2_traffic.json profiles/r2_traffic.json
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2_final_bench.log 2> gpurun_out/r2_final_bench.err; echo "bench rc=$? elapsed $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_ref.log 2> gpurun_out/r2_final_bench_ref.err; echo "ref rc=$? elapsed $(( $(date +%s) - t0 )) s"; cut -c1-400 gpurun_out/r2_final_bench_ref.log
python - <<'PY'
import json
for l in open('gpurun_out/r2_final_bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('roofline_split'), d['vs_gpu_incumbent'], d['cpu_baseline'].get('value'), d['extra']['c3']['value'], d['extra']['dropout_0'], d['clocks'])
PY
