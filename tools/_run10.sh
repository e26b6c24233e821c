set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --check-ddp-shapes > gpurun_out/r2_bench_2gpu.log 2>&1; echo "rc=$?"; tail -c 3000 gpurun_out/r2_bench_2gpu.log
