for v in _prev "" _prev ""; do
  PBRT_B200_LIB=pbrt-rust_b200/libpbrt_b200$v.so python bench.py --no-cpu-baseline --no-e2e --steps 6 --warmup 3 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', 'roofline', round(d['roofline']['frac'],3), {k: round(v,1) for k,v in d['kernel_ms'].items()})"
done
