# round 2, GPU call D: warp-cooperative traversal (full-mask votes), with / without the postponed leaf, against the per-lane form
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
for v in _nc "" _cp; do
  export PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200$v.so
  echo "=== variant '$v'"
  python tools/trace_ab3.py 2>&1 | tail -12
  python tools/step_diag.py 2>&1 | grep -E "plain"
done > gpurun_out/r2d_ab.log 2>&1
cat gpurun_out/r2d_ab.log
