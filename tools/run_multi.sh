# multi-GPU bench lines (run with: gpurun --gpus 8 -- 'bash tools/run_multi.sh 8')
mkdir -p gpurun_out
for n in "$@"; do
extra=""; [ "$n" = "2" ] && extra="--check-ddp-shapes"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 $extra > gpurun_out/r2_bench_${n}gpu.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/r2_bench_${n}gpu.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), d.get('ddp_check'))
"
done
