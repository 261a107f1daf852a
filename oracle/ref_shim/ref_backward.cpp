// oracle/ref_shim/ref_backward.cpp — TEST INFRASTRUCTURE (see optix.h in this directory).
// Compiles the reference's backward device program (DLT/optix_tracer/backward.cu) unmodified as
// host code and exposes the equivalent of TraceSurfelsBackwardCUDA
// (DLT/trace_surfels.cpp:268-386) as a flat C function. Built only into oracle/_ref/.
#include "optix.h"
#include "backward.cu"       // found via -I <reference>/submodules/diff-lidar-tracer/optix_tracer
#include "shim_trace.inl"
#include <cstring>

extern "C" int ref_backward(int H, int W, int P, int D, int M,
                            const float* ray_o, const float* ray_d, const float* vertices,
                            const float* bg, const float* means, const float* shs,
                            const float* opac, const float* scales, float scale_modifier,
                            const float* rots, const float* out_f32, const float* dL_dout,
                            float* dL_dmeans, float* dL_dshs, float* dL_dopac,
                            float* dL_dscales, float* dL_drots)
{
    // trace_surfels.cpp:322-329: zero-initialised gradient tensors
    memset(dL_dmeans, 0, sizeof(float) * (size_t)P * 3);
    memset(dL_dshs, 0, sizeof(float) * (size_t)P * M * 3);
    memset(dL_dopac, 0, sizeof(float) * (size_t)P);
    memset(dL_dscales, 0, sizeof(float) * (size_t)P * 2);
    memset(dL_drots, 0, sizeof(float) * (size_t)P * 4);
    memset(&params, 0, sizeof(params));
    params.P = P; params.H = H; params.W = W; params.D = D; params.M = M;
    params.ray_o = (float3*)ray_o; params.ray_d = (float3*)ray_d;
    params.vertices = (float3*)vertices;
    params.background = (float*)bg;
    params.means3D = (glm::vec3*)means;
    params.shs = (float*)shs;
    params.colors_precomp = nullptr;
    params.opacities = (float*)opac;
    params.scales = (glm::vec2*)scales;
    params.scale_modifier = scale_modifier;
    params.rotations = (glm::vec4*)rots;
    params.out_attr_float32 = (float*)out_f32;
    params.dL_dout_attr_float32 = (float*)dL_dout;
    params.dL_dmeans3D = (glm::vec3*)dL_dmeans;
    params.dL_dshs = (glm::vec3*)dL_dshs;
    params.dL_dopacities = dL_dopac;
    params.dL_dscales = (glm::vec2*)dL_dscales;
    params.dL_drotations = (glm::vec4*)dL_drots;
    shim_launch(H, W);
    return 0;
}
