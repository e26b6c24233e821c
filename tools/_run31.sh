mkdir -p gpurun_out /tmp/ncu
i=0
for spec in "gemm_tc_kernel<\(int\)2:0" "gemm_tc_kernel<\(int\)3:4" "gemm_tc_kernel<\(int\)4:10" "layernorm_bwd_kernel<__nv_bfloat16:30" "attn_pair_kernel:2" "ffn_bwd_kernel:2" "rel_attn_fwd_kernel:2" "attn_softmax_bwd_kernel:8" "glu_dwconv_fwd_stream_kernel:2" "dwconv_glu_bwd_kernel:2"; do
k="${spec%%:*}"; s="${spec##*:}"; i=$((i+1))
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:"$k" -s $s -c 1 -f -o /tmp/ncu/k$i python tools/one_step.py c2 bf16 0.1 > /tmp/ncu/log$i.txt 2>&1
ncu -i /tmp/ncu/k$i.ncu-rep --page raw --csv > gpurun_out/r2q_k$i.csv 2>/dev/null
done
ls -la gpurun_out/r2q_*.csv
