mkdir -p gpurun_out
for g in 5 8 10 12 15 20; do echo "== PBRT_B200_GRID_SHADE=$g"; PBRT_B200_GRID_SHADE=$g python tools/step_diag.py 2>&1 | grep -E "plain"; done > gpurun_out/r2o_grid.log 2>&1
cat gpurun_out/r2o_grid.log
