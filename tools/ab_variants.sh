# usage: bash tools/ab_variants.sh "" _prev ...   (suffixes of pbrt-rust_b200/libpbrt_b200<suffix>.so); trace batches + whole-step timing
for v in "$@"; do
  export PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200$v.so
  echo "=== variant '$v'"
  if [ -z "$SKIP_TRACE" ]; then AB_QUICK=2 python tools/trace_ab3.py 2>&1 | tail -2; fi
  python tools/step_diag.py 2>&1 | grep -E "plain"
done
