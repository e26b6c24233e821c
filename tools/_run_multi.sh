# multi-GPU bench lines (run with: gpurun --gpus 8 -- 'bash tools/_run_multi.sh 8 4')
mkdir -p gpurun_out
for n in "$@"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_${n}gpu.log 2>&1; echo "rc=$?"
done
