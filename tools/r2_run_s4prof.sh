mkdir -p gpurun_out
export DIAG_SCENE=s4 DIAG_SPP=4
python tools/step_diag.py 2>&1 | grep plain > gpurun_out/r2s4_step.log; cat gpurun_out/r2s4_step.log
ncu --set full --import-source on --clock-control none --kernel-name "regex:k_trace_closest" --launch-skip 12 --launch-count 2 -o /tmp/s4c -f python tools/step_diag.py > gpurun_out/r2s4_closest.log 2>&1
ncu -i /tmp/s4c.ncu-rep --page raw --csv > gpurun_out/r2s4_closest_raw.csv 2>/dev/null
for k in 0 1; do ncu -i /tmp/s4c.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $k --launch-count 1 2>/dev/null | gzip > gpurun_out/r2s4_closest_src_$k.csv.gz; done
du -sh gpurun_out
