for p in 33554432 67108864 134217728; do python bench.py --no-cpu-baseline --no-e2e --steps 4 --warmup 3 --paths-in-flight $p 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print($p, round(d['value']/1e6,1), 'Msamples/s', round(d['ms_per_step'],2), 'ms')"; done
