# round 2, GPU call E: parity of the current build, step breakdown vs the previous best (_nc), one full bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -3 gpurun_out/r2e_pytest.log
SKIP_TRACE=1 bash tools/ab_variants.sh _nc "" > gpurun_out/r2e_ab.log 2>&1
cat gpurun_out/r2e_ab.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -c 2500 gpurun_out/r2e_bench.json
