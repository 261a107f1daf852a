// oracle/ref_shim/optix.h — TEST INFRASTRUCTURE, never shipped, never on the product path.
//
// Host-side stand-in for the parts of the OptiX 7.7 *device* API that the reference tracer's
// device programs use (forward.cu:48-63,146-356, backward.cu:49-64,434-739). With this header
// first on the include path, the reference's own forward.cu / backward.cu are compiled
// UNMODIFIED, from where they lie under /root/reference, as plain host C++ (see
// oracle/build_ref.sh). Only the closed-source pieces are replaced:
//   * optixTrace + the GAS  -> a loop over the 2P proxy triangles of build2DRectangle
//                              (primitive_utils.py:182-224) with a double-precision
//                              Moeller-Trumbore test, invoking the reference's own
//                              __anyhit__ot for every hit with tmin < t < tmax
//   * optixLaunch           -> an OpenMP loop over the (H, W) launch grid
// Everything else (k-buffer, round loop, compositing, VJPs, atomics) is the reference's code.
#pragma once

#include <cuda_runtime.h>   // float3/make_float3/dim3/uint3 and empty __device__/__global__ on host
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <algorithm>

typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
enum { OPTIX_RAY_FLAG_NONE = 0 };

// CUDA device built-ins the reference headers use that a host compiler does not have.
static inline float min(float a, float b) { return a < b ? a : b; }
static inline float max(float a, float b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(unsigned a, int b) { return (int)a < b ? (int)a : b; }
static inline int max(int a, unsigned b) { return a > (int)b ? a : (int)b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }   // IEEE stand-in for the approximate intrinsic
static inline void __trap() { abort(); }
static inline float atomicAdd(float* addr, float v)
{
    float old;
#pragma omp atomic capture
    { old = *addr; *addr += v; }
    return old;
}

// Per-thread launch / traversal state of the stand-in.
struct ShimThreadState {
    uint3 launch_index;
    uint3 launch_dims;
    unsigned int payload0, payload1;
    float cur_tmax;
    unsigned int cur_prim;
};
extern thread_local ShimThreadState shim_ts;

static inline uint3 optixGetLaunchIndex() { return shim_ts.launch_index; }
static inline uint3 optixGetLaunchDimensions() { return shim_ts.launch_dims; }
static inline unsigned int optixGetPayload_0() { return shim_ts.payload0; }
static inline unsigned int optixGetPayload_1() { return shim_ts.payload1; }
static inline float optixGetRayTmax() { return shim_ts.cur_tmax; }
static inline unsigned int optixGetPrimitiveIndex() { return shim_ts.cur_prim; }
static inline void optixIgnoreIntersection() {}

// Defined in shim_trace.inl (included after the reference source, because it needs `params`).
void optixTrace(OptixTraversableHandle handle, float3 ray_o, float3 ray_d, float tmin, float tmax,
                float ray_time, OptixVisibilityMask mask, unsigned int flags,
                unsigned int sbt_offset, unsigned int sbt_stride, unsigned int miss_index,
                unsigned int& p0, unsigned int& p1);
