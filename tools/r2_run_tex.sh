# ncu --set full of the textured shade kernel (k_shade<Q_TEX>) on T1 at 1920x1080, 4 spp: the first launch (camera vertices: EWA lookups
# with real footprints) and the third (bounce 2: zero differentials).  CSV exports only.
TAG=${1:-r2tex}
export DIAG_SCENE=t1 DIAG_SPP=4
cap() {
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name "regex:$2" --launch-skip $3 --launch-count $4 -o /tmp/${TAG}_$1 -f python tools/step_diag.py > gpurun_out/${TAG}_$1.log 2>&1
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
  for ((k = 0; k < $4; k++)); do ncu -i /tmp/${TAG}_$1.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $k --launch-count 1 2>/dev/null | gzip > gpurun_out/${TAG}_$1_src_$k.csv.gz; done
}
cap shade 'k_shade<.{0,8}7,' 0 3
tail -5 gpurun_out/${TAG}_shade.log
