set -x
mkdir -p gpurun_out
python tools/ffn_bench.py 37674 256 2048 1 > gpurun_out/r2_ffn_bench_2.log 2>&1
python tools/ffn_bench.py 37674 256 2048 0 >> gpurun_out/r2_ffn_bench_2.log 2>&1
cat gpurun_out/r2_ffn_bench_2.log
ncu --set full --clock-control none --import-source on -k regex:ffn_bwd_kernel --launch-skip 2 -c 1 -o gpurun_out/r2_ffn_bwd_v1 python tools/ffn_bench.py 37674 256 2048 1 fused > gpurun_out/r2_ncu_ffn.log 2>&1; tail -3 gpurun_out/r2_ncu_ffn.log
