mkdir -p gpurun_out
export DIAG_SPP=64
( echo "== default"; python tools/step_diag.py 2>&1 | grep -E "plain:"
for r in 16 20 28 32; do echo "== PBRT_B200_REFILL=$r"; PBRT_B200_REFILL=$r python tools/step_diag.py 2>&1 | grep -E "plain:"; done
for m in 0 8 12 20 24; do echo "== PBRT_B200_INTERIOR_MIN=$m"; PBRT_B200_INTERIOR_MIN=$m python tools/step_diag.py 2>&1 | grep -E "plain:"; done
for p in 33554432 134217728; do echo "== paths_in_flight=$p"; python tools/step_diag.py $p 2>&1 | grep -E "plain:"; done ) > gpurun_out/r2q_tune.log 2>&1
cat gpurun_out/r2q_tune.log
