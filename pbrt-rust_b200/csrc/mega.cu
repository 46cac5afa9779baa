// Third translation unit of the path tracer: the one-kernel forms k_zt_mega ((0,2)-sequence sampler under the path integrator) and
// k_vol_mega (volpath), compiled from the SAME source as render.o with the same exact-arithmetic flags.  Their out-of-line per-material
// shade bodies are the slowest thing ptxas sees in this library; as a unit of their own they compile next to render.o instead of after it.
#define PB_TU_MEGA 1
#include "render.cu"
