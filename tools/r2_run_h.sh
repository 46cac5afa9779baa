mkdir -p gpurun_out
bash tools/ab_variants.sh "" _pf > gpurun_out/r2h_ab.log 2>&1
cat gpurun_out/r2h_ab.log
bash tools/profile_pass.sh r2h > gpurun_out/r2h_profile.log 2>&1
tail -3 gpurun_out/r2h_profile.log
