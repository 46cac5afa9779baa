mkdir -p gpurun_out
export DIAG_SPP=64
( for v in _rf24 ""; do echo "== variant '$v'"; PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200$v.so python tools/step_diag.py 2>&1 | grep -E "plain"; PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200$v.so python tools/step_diag.py 134217728 2>&1 | grep -E "plain:"; done ) > gpurun_out/r2r_refill.log 2>&1
cat gpurun_out/r2r_refill.log
