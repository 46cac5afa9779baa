# A/B of the compressed quad nodes (PB_CQUAD=1, libpbrt_b200_cq.so) against the default build: ray batches, hit-record differences, S3 frame
mkdir -p gpurun_out /tmp/cqd
( for v in "" _cq "" _cq; do echo "== variant '$v'"; AB_QUICK=2 AB_DUMP=/tmp/cqd/h$v PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200$v.so python tools/trace_ab3.py 2>&1 | grep -E "default|lib"; done
  python tools/r2_cq_diff.py /tmp/cqd/h /tmp/cqd/h_cq
  for v in "" _cq "" _cq; do echo "== variant '$v'"; DIAG_SPP=16 PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200$v.so python tools/step_diag.py 2>&1 | grep -E "plain:"; done
  PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200_cq.so timeout 600 python -m pytest tests/test_gpu_intersect.py tests/test_gpu_render.py -m gpu -q 2>&1 | tail -5
) > gpurun_out/r2_cq.log 2>&1
cat gpurun_out/r2_cq.log
