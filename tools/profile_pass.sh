# One profiling pass for profiles/ (run under gpurun): ncu launch list of a bench step, and ncu --set full of the LARGEST launches of a
# steady-state 16-spp S3 step (iteration 0 of the third render call of tools/step_diag.py: 12 wavefront iterations per call).
# Only CSV exports are kept: the .ncu-rep files together exceed gpurun's 64 MiB return limit.
TAG=${1:-r02}
mkdir -p gpurun_out
export DIAG_SPP=16
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches.log 2>&1
cap() {  # name, kernel regex, launches to skip, launches to take
  ncu --set full --import-source on --clock-control none --kernel-name "regex:$2" --launch-skip $3 --launch-count $4 -o /tmp/${TAG}_$1 -f python tools/step_diag.py > gpurun_out/${TAG}_$1.log 2>&1
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
  for ((k = 0; k < $4; k++)); do ncu -i /tmp/${TAG}_$1.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $k --launch-count 1 2>/dev/null | gzip > gpurun_out/${TAG}_$1_src_$k.csv.gz; done
}
cap closest 'k_trace_closest' 24 2
cap shadow 'k_trace_shadow' 24 1
cap shade 'k_shade' 168 7
cap regen 'k_finish_regen' 26 2
cap classify 'k_classify' 24 1
du -sh gpurun_out
