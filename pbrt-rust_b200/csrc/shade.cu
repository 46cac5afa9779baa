// Second translation unit of the wavefront path tracer: the shade kernels (k_shade<material>, k_rec_shade) and their
// launchers, compiled from the SAME source as render.o but with --use_fast_math (see the note at the top of render.cu
// and csrc/Makefile).  Nothing here is pinned bit for bit: BSDF / light / material arithmetic is tolerance-parity
// (image relMSE <= 1e-3 vs the oracle, tests/test_gpu_render.py), while every ray the shade kernels emit is still
// traced by the exactly-rounded kernels of render.o.
#define PB_TU_SHADE 1
#include "render.cu"
