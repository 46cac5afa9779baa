"""Where does a bench step's time go?  Per-phase CUDA-event times of pbrt_b200_render on S3, with and without the
nvidia-smi clock sampler running, and with stats (timing events) on/off.  Run under gpurun."""
import importlib, subprocess, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
P = importlib.import_module("pbrt-rust_b200")
import os
SCENE = os.environ.get("DIAG_SCENE", "s3")
SPP = int(os.environ.get("DIAG_SPP", "16"))
setup = {"s3": P.scenes.displaced_sphere_scene, "s4": P.scenes.foliage_field_scene, "cornell": P.scenes.cornell_scene,
         "s5": lambda: P.scenes.glass_knot_scene(nu=4096, nv=640), "t1": lambda: P.scenes.textured_scene(xres=1920, yres=1080)}[SCENE]()
integ = setup.make_integrator(spp_=SPP * 16)
film = integ.film
sc = P.Scene(setup.flat)
film_t = torch.zeros((film.width * film.height, 4), dtype=torch.float32, device="cuda")
pif = int(sys.argv[1]) if len(sys.argv) > 1 else 0


def run(tag, n=4, k0=0):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    agg = None
    for k in range(n):
        film_t.zero_()
        _, st = sc.render(integ, sample_range=((k0 + k) * SPP, (k0 + k + 1) * SPP), device_ptr=film_t.data_ptr(), paths_in_flight=pif)
        agg = st
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    print(f"{tag}: {e0.elapsed_time(e1) / n:.1f} ms/step (wall {wall:.1f}); last step: device {agg.device_ms:.1f} closest {agg.trace_closest_ms:.1f} "
          f"shadow {agg.trace_any_ms:.1f} shade {agg.shade_ms:.1f} finish {agg.finish_ms:.1f} iters {agg.iterations} launches {agg.kernel_launches} "
          f"-> {agg.camera_rays / agg.device_ms / 1e3:.1f} Msamples/s", flush=True)


run("warmup", 2)
run("plain", 4, 2)
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
time.sleep(0.5)
run("with nvidia-smi -lms 100", 4, 6)
proc.terminate()
print(proc.stdout.read()[-400:])
run("plain again", 4, 10)
sc.close()
