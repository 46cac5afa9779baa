"""Small textured + volumetric renders for compute-sanitizer (memcheck / racecheck / synccheck): every new kernel path of the round --
k_shade<Q_TEX> behind its whole-queue counting sort (k_tex_count, k_tex_scan, k_tex_scatter), k_rec_shade<.., TEX> (two-candidate frames, differentials), k_zt_mega and
k_vol_mega with the textured case, k_vol_mega on the fog box."""
import importlib, sys
sys.path.insert(0, '.')
P = importlib.import_module("pbrt-rust_b200")
S = P.scenes
for sampler, integ in (("sobol", "path"), ("halton", "whitted"), ("sobol", "directlighting:all"), ("02sequence", "path"), ("02sequence", "whitted"), ("sobol", "volpath")):
    setup = S.textured_scene(xres=48, yres=32, spp=2, sampler=sampler)
    sc = P.Scene(setup.flat)
    img, st = sc.render(setup.make_integrator(integrator=integ))
    sc.close()
    print(sampler, integ, float(img.mean()), st.kernel_launches, flush=True)
setup = S.fog_box_scene(xres=32, yres=32, spp=2)
sc = P.Scene(setup.flat)
img, st = sc.render(setup.make_integrator())
sc.close()
print("fog", float(img.mean()), flush=True)
