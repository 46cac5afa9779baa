# round 2, GPU call G: parity, then grid size of the streaming kernels (finish/regen, classify), then a bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
for g in 4 6 8 12; do echo "== PBRT_B200_GRID_SMALL=$g"; PBRT_B200_GRID_SMALL=$g python tools/step_diag.py 2>&1 | grep -E "plain"; done > gpurun_out/r2g_grid.log 2>&1
cat gpurun_out/r2g_grid.log
