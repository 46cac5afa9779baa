mkdir -p gpurun_out
bash tools/ab_variants.sh _fg1 "" _fg8 > gpurun_out/r2j_ab.log 2>&1
cat gpurun_out/r2j_ab.log
python bench.py --scene s4 --steps 2 --warmup 1 --spp 8 --no-cpu-baseline --no-e2e > gpurun_out/r2j_s4.json 2> gpurun_out/r2j_s4.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_s4.json')); print('S4 value',d['value']/1e6,'ms',d['ms_per_step'],d['kernel_ms'])"
