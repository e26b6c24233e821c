mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_u2_gpu.py -q -x --timeout=250 2>&1 | tail -2
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2_final_bench.log 2> gpurun_out/r2_final_bench.err; echo "bench rc=$? elapsed $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
for l in open('gpurun_out/r2_final_bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['vs_gpu_incumbent'], d['cpu_baseline'].get('value'), d['extra']['c3']['value'], d['clocks'])
        print(d['extra']['variable_shapes'])
PY
