set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2e_step_traffic.csv python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2e_ncu.log 2>&1; echo "ncu rc=$?"
for k in glu_dwconv_fwd_kernel dwconv_glu_bwd_kernel layernorm_bwd_kernel attn_softmax_bwd_kernel; do
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/r2e_$k python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2e_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -s 30 -c 12 -f -o gpurun_out/r2e_gemm12 python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2e_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
