mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -3 gpurun_out/r2m_pytest.log
python tools/step_diag.py 2>&1 | grep -E "plain" > gpurun_out/r2m_step.log; cat gpurun_out/r2m_step.log
DIAG_SCENE=s4 DIAG_SPP=4 python tools/step_diag.py 2>&1 | grep -E "plain" > gpurun_out/r2m_step_s4.log; cat gpurun_out/r2m_step_s4.log
