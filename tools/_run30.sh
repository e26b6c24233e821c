mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_pytest.log; tail -3 gpurun_out/r2p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2p_smoke.log
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"gemm_tc_kernel|layernorm_bwd_kernel|attn_pair_kernel|ffn_bwd_kernel" -s 120 -c 40 -f -o gpurun_out/r2p_mix python tools/one_step.py c2 bf16 0.1 > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
