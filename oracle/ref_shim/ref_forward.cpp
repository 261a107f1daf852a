// oracle/ref_shim/ref_forward.cpp — TEST INFRASTRUCTURE (see optix.h in this directory).
// Compiles the reference's forward device program (DLT/optix_tracer/forward.cu) unmodified as
// host code and exposes the equivalent of TraceSurfelsCUDA (DLT/trace_surfels.cpp:151-265) as a
// flat C function. Built only into oracle/_ref/ by oracle/build_ref.sh.
#include "optix.h"
#include "forward.cu"        // found via -I <reference>/submodules/diff-lidar-tracer/optix_tracer
#include "shim_trace.inl"
#include <cstring>

extern "C" int ref_forward(int H, int W, int P, int D, int M,
                           const float* ray_o, const float* ray_d, const float* vertices,
                           const float* bg, const float* means, const float* shs,
                           const float* opac, const float* scales, float scale_modifier,
                           const float* rots, float* out_f32, float* accum_w)
{
    // trace_surfels.cpp:207-209: zeros (H,W,9), zeros (P)
    memset(out_f32, 0, sizeof(float) * (size_t)H * W * NUM_CHANNELS_F);
    memset(accum_w, 0, sizeof(float) * (size_t)P);
    memset(&params, 0, sizeof(params));
    params.P = P; params.H = H; params.W = W; params.D = D; params.M = M;   // :215-225
    params.training = true;
    params.ray_o = (float3*)ray_o; params.ray_d = (float3*)ray_d;
    params.vertices = (float3*)vertices;
    params.background = (float*)bg;
    params.means3D = (glm::vec3*)means;
    params.shs = (float*)shs;
    params.colors_precomp = nullptr;          // empty tensor -> null data_ptr (:233)
    params.opacities = (float*)opac;
    params.scales = (glm::vec2*)scales;
    params.scale_modifier = scale_modifier;
    params.rotations = (glm::vec4*)rots;
    params.out_attr_float32 = out_f32;
    params.out_attr_uint32 = nullptr;
    params.accum_gaussian_weights = accum_w;
    shim_launch(H, W);                        // optixLaunch(..., H, W, 1) :256
    return 0;
}
