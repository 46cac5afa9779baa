# round 2, GPU call F: shared-memory traversal stack (first N levels in shared memory, the rest in local memory): N = 0 / 4 / 8 / 16
mkdir -p gpurun_out
bash tools/ab_variants.sh _sh0 _sh4 "" _sh16 > gpurun_out/r2f_ab.log 2>&1
cat gpurun_out/r2f_ab.log
