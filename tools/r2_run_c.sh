# round 2, GPU call C: parity + A/B of the two-translation-unit build (fast-math shade.o), ncu of the large launches (CSV exports only: the
# .ncu-rep files together exceed gpurun's 64 MiB return limit)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
bash tools/ab_variants.sh _t1 _fm _pl "" > gpurun_out/r2c_ab.log 2>&1
cat gpurun_out/r2c_ab.log
export DIAG_SPP=16
ncu --set full --import-source on --clock-control none --kernel-name 'regex:k_shade' --launch-skip 182 --launch-count 7 -o /tmp/r2c_shade -f python tools/step_diag.py > gpurun_out/r2c_shade.log 2>&1
ncu -i /tmp/r2c_shade.ncu-rep --page raw --csv > gpurun_out/r2c_shade_raw.csv 2>/dev/null
for k in 1 2 5; do ncu -i /tmp/r2c_shade.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $k --launch-count 1 2>/dev/null | gzip > gpurun_out/r2c_shade_src_$k.csv.gz; done
ncu --set full --import-source on --clock-control none --kernel-name 'regex:k_trace_closest|k_classify|k_trace_shadow|k_trace_mis|k_finish_regen' --launch-skip 130 --launch-count 5 -o /tmp/r2c_trace -f python tools/step_diag.py > gpurun_out/r2c_trace.log 2>&1
ncu -i /tmp/r2c_trace.ncu-rep --page raw --csv > gpurun_out/r2c_trace_raw.csv 2>/dev/null
for k in 0 1 2 3 4; do ncu -i /tmp/r2c_trace.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $k --launch-count 1 2>/dev/null | gzip > gpurun_out/r2c_trace_src_$k.csv.gz; done
du -sh gpurun_out
