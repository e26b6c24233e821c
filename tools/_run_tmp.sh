mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_shape_gpu.py tests/test_conv_planes_gpu.py -x -q -s --timeout=800 2>&1 | grep -v Warning | grep -v "^$" | tail -25
t0=$(date +%s)
timeout 1500 python bench.py --batch 252 > gpurun_out/r2z_bench252.log 2> gpurun_out/r2z_bench252.err; echo "bench rc=$? elapsed $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
for l in open('gpurun_out/r2z_bench252.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['gpu_incumbent'], d['cpu_baseline'], d['extra']['c3'])
PY
