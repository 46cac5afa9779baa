# One profiling pass for profiles/: bench line, ncu launch list of the same command, ncu --set full of one steady-state iteration.
TAG=${1:-s8}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name 'regex:k_trace_closest|k_shade|k_finish_regen|k_trace_shadow|k_classify' --launch-skip 24 --launch-count 12 -o gpurun_out/${TAG}_full -f python tools/step_diag.py > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
tail -c 1500 gpurun_out/${TAG}_bench.json
