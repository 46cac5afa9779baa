mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -3 gpurun_out/r2k_pytest.log
bash tools/ab_variants.sh _fg1 "" > gpurun_out/r2k_ab.log 2>&1
cat gpurun_out/r2k_ab.log
