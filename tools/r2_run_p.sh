mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -3 gpurun_out/r2p_pytest.log
SKIP_TRACE=1 bash tools/ab_variants.sh _so "" > gpurun_out/r2p_ab.log 2>&1
cat gpurun_out/r2p_ab.log
for v in _so ""; do PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200$v.so python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,1), 'Msamples/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']/1e6,1), {k: round(v,1) for k,v in d['kernel_ms'].items()})"; done > gpurun_out/r2p_bench.log 2>&1
cat gpurun_out/r2p_bench.log
