# round 2, GPU call B: ncu --set full of the FIRST (largest) iteration of a steady-state step: every wavefront kernel once
mkdir -p gpurun_out
export DIAG_SPP=16
# a 16-spp step of step_diag = one drained wave: iteration 0 of step k starts after k * (13 iterations * 14 launches + 3).  Skip the two warm-up steps by name-filtered counts:
# k_trace_closest launches per step = 13 -> skip 26 (2 warm-up steps), take the first of the third step; same for the others.
ncu --set full --import-source on --clock-control none --kernel-name 'regex:k_trace_closest|k_classify|k_trace_shadow|k_trace_mis|k_finish_regen' --launch-skip 130 --launch-count 5 -o gpurun_out/r2b_trace -f python tools/step_diag.py > gpurun_out/r2b_trace.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name 'regex:k_shade' --launch-skip 182 --launch-count 7 -o gpurun_out/r2b_shade -f python tools/step_diag.py > gpurun_out/r2b_shade.log 2>&1
ncu -i gpurun_out/r2b_trace.ncu-rep --page raw --csv > gpurun_out/r2b_trace_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_shade.ncu-rep --page raw --csv > gpurun_out/r2b_shade_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail
