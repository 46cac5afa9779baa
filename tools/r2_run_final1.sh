# parity + bench + profile pass of the (near-)final build
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2f1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f1_pytest.log
tail -3 gpurun_out/r2f1_pytest.log
python bench.py > gpurun_out/r2f1_bench.json 2> gpurun_out/r2f1_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2f1_bench.json')); print('value',d['value']/1e6,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e6,d['kernel_ms'],d['ray_batches'], 'cpu', d['cpu_baseline']['value']/1e6)"
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2f1_ref.json 2> gpurun_out/r2f1_ref.err; cut -c1-400 gpurun_out/r2f1_ref.json
bash tools/profile_pass.sh r2f1 > gpurun_out/r2f1_profile.log 2>&1
tail -2 gpurun_out/r2f1_profile.log
