// placeholder until the wavefront lands
#include "scene.cuh"
#include "util.cuh"
namespace pb { void render_release_scene_state(pbrt_b200_scene*) {} }
extern "C" int pbrt_b200_render(pbrt_b200_scene*, const pbrt_b200_render_desc*, float*, pbrt_b200_render_stats*) {
    return fail(PBRT_B200_ERR_UNSUPPORTED, "render: not built yet");
}
