mkdir -p gpurun_out
timeout 300 python tools/step_profile.py c2 0 40 0 > gpurun_out/r2_step_profile_p0.log 2>&1
timeout 300 python tools/step_profile.py c2 0 40 0.1 > gpurun_out/r2_step_profile_p1.log 2>&1
