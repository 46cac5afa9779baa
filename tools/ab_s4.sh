python -m pytest tests/test_gpu_instancing.py tests/test_gpu_intersect.py tests/test_gpu_render.py -m gpu -x -q 2>&1 | tail -3
for v in _prev ""; do echo "== $v"; PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200$v.so DIAG_SCENE=s4 DIAG_SPP=4 python tools/step_diag.py 2>&1 | grep -E "plain"; done
