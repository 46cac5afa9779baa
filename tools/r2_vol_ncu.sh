# ncu --set full of k_vol_mega on the fog box (512x512, 8 spp): counters + top source lines.  CSV exports only.
mkdir -p gpurun_out
cat > /tmp/vol_one.py <<'P'
import importlib, sys
sys.path.insert(0, '.')
P = importlib.import_module("pbrt-rust_b200")
fog = P.scenes.fog_box_scene(xres=512, yres=512, spp=8)
sc = P.Scene(fog.flat)
img, st = sc.render(fog.make_integrator(spp_=8))
print(st.camera_rays, st.device_ms)
P
ncu --set full --import-source on --clock-control none --kernel-name "regex:k_vol_mega" --launch-count 1 -o /tmp/r2vol -f python /tmp/vol_one.py > gpurun_out/r2vol_ncu.log 2>&1
ncu -i /tmp/r2vol.ncu-rep --page raw --csv > gpurun_out/r2vol_raw.csv 2>/dev/null
ncu -i /tmp/r2vol.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/r2vol_src.csv.gz
tail -3 gpurun_out/r2vol_ncu.log
