set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q --timeout=300 > gpurun_out/r2i_kernels.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2i_kernels.log
for k in glu_dwconv_fwd_stream_kernel dwconv_glu_bwd_stream_kernel attn_pair_kernel ffn_fwd_kernel; do
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/r2i_$k python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2i_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
(time timeout 600 python bench.py --steps 20 --warmup 5 --quick) > gpurun_out/r2i_bench_quick.log 2>&1
LASR_FUSED_FFN_FWD=0 timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2i_bench_quick_nofwd.log 2>&1
LASR_FUSED_FFN_FWD=0 LASR_CONVMOD_STREAM=0 timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2i_bench_quick_nofwd_nostream.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r2i_bench_quick.log', 'gpurun_out/r2i_bench_quick_nofwd.log', 'gpurun_out/r2i_bench_quick_nofwd_nostream.log'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f, d['value'], d['ms_per_step'], d.get('e2e', {}).get('ms_per_step'), d.get('roofline', {}).get('frac'))
PY
