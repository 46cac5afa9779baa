#!/usr/bin/env python3
"""pbrt-rust's `main` for the B200 path (src/main.rs): parse scene files and render them.

    python tools/render_file.py scene.pbrt [more.pbrt ...] [--outfile out.pfm] [--device 0] [--quick] [--cropwindow x0 x1 y0 y1]

Flags mirror the reference CLI's (--outfile, --quick, --cropwindow; README.md:16-40).  The image is written as PFM (or .npy);
other extensions named by the scene's Film are replaced by .pfm (image encoders are out of scope).  Needs a CUDA device:
there is no CPU fallback."""
import argparse
import importlib
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scenes", nargs="+")
    ap.add_argument("--outfile", default="")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--cropwindow", type=float, nargs=4, metavar=("X0", "X1", "Y0", "Y1"))
    a = ap.parse_args()
    pkg = importlib.import_module("pbrt-rust_b200")
    crop = ((a.cropwindow[0], a.cropwindow[1]), (a.cropwindow[2], a.cropwindow[3])) if a.cropwindow else ((0.0, 1.0), (0.0, 1.0))
    for path in a.scenes:
        t0 = time.time()
        api = pkg.pbrt_parse(path, quick_render=a.quick, image_file=a.outfile, crop_window=crop)
        t1 = time.time()
        for job in api.jobs:
            img, stats = job.render(device=a.device)
            out = job.write_image(job.filename, img)
            t2 = time.time()
            print(f"{path}: parsed in {t1 - t0:.2f} s, rendered {stats.camera_rays} camera samples in {stats.device_ms:.1f} ms on device {a.device} "
                  f"({stats.camera_rays / max(stats.device_ms, 1e-9) / 1e3:.1f} M samples/s; {t2 - t1:.2f} s with upload and download) -> {out}")
            t1 = t2


if __name__ == "__main__":
    main()
