mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -3 gpurun_out/r2i_pytest.log
python tools/step_diag.py 2>&1 | grep -E "plain" > gpurun_out/r2i_step.log; cat gpurun_out/r2i_step.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); print('value',d['value']/1e6,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e6,d['kernel_ms'],d['ray_batches'])"
