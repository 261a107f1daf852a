// lrt_rays.cu — LiDAR range-image ray generation and range -> point back-projection, one kernel each.
//
// SURVEY.md §8(f) N3. Replaces LiDARSensor.get_range_rays (lib/scene/lidar_sensor.py:395-434) and
// LiDARSensor.range2point (:325-393) of the reference (~15 torch kernels over (H, W) grids each):
//   azimuth(w)      = ((W - w) - pixel_offset) / W * 2 pi - pi - angle_offset
//   inclination(h)  = table[H-1-h]                                  (Waymo: a list of H beam inclinations, ascending)
//                   = ((H - h) - pixel_offset) / H * (hi - lo) + lo  (KITTI: two bounds)
//   d_sensor        = (cos i cos a, cos i sin a, sin i)
//   rays:   d = normalize(R d_sensor), origin = t            (R | t = sensor2world[:3])
//   points: p = R (normalize(d_sensor) * range) + t
// The shared origin is returned once (3 floats): the tracer takes it with ray_o_stride = 0, which is what the
// reference's expanded stride-0 view amounts to.
#include "lrt_ctx.cuh"

namespace {

struct RayGrid { int H, W; const float* inc_table; float inc_lo, inc_hi, pixel_offset, angle_offset; const float* s2w; };

__device__ __forceinline__ void sensor_dir(const RayGrid& g, int h, int w, float* d)
{
    const float pi = 3.14159265358979323846f;
    const float x = ((float)(g.W - w) - g.pixel_offset) / (float)g.W;
    const float az = x * 2.0f * pi - pi - g.angle_offset;
    float inc;
    if (g.inc_table) inc = g.inc_table[g.H - 1 - h];
    else inc = ((float)(g.H - h) - g.pixel_offset) / (float)g.H * (g.inc_hi - g.inc_lo) + g.inc_lo;
    const float ci = cosf(inc);
    d[0] = ci * cosf(az); d[1] = ci * sinf(az); d[2] = sinf(inc);
}

template <bool POINTS>
__global__ void __launch_bounds__(256) k_range_rays(RayGrid g, const float* __restrict__ range_map, float* __restrict__ out, float* __restrict__ centre)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && centre) { centre[0] = g.s2w[3]; centre[1] = g.s2w[7]; centre[2] = g.s2w[11]; }
    if (i >= g.H * g.W) return;
    const int h = i / g.W, w = i - h * g.W;
    float d[3];
    sensor_dir(g, h, w, d);
    float R[9];
#pragma unroll
    for (int r = 0; r < 3; r++) { R[3 * r] = g.s2w[4 * r]; R[3 * r + 1] = g.s2w[4 * r + 1]; R[3 * r + 2] = g.s2w[4 * r + 2]; }
    float o[3];
    if (POINTS) {
        const float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const float rg = range_map[i];
        const float p[3] = {d[0] / n * rg, d[1] / n * rg, d[2] / n * rg};
#pragma unroll
        for (int r = 0; r < 3; r++) o[r] = (p[0] * R[3 * r] + p[1] * R[3 * r + 1] + p[2] * R[3 * r + 2]) + g.s2w[4 * r + 3];
    } else {
        float v[3];
#pragma unroll
        for (int r = 0; r < 3; r++) v[r] = d[0] * R[3 * r] + d[1] * R[3 * r + 1] + d[2] * R[3 * r + 2];
        const float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        o[0] = v[0] / n; o[1] = v[1] / n; o[2] = v[2] / n;
    }
    out[3 * (size_t)i] = o[0]; out[3 * (size_t)i + 1] = o[1]; out[3 * (size_t)i + 2] = o[2];
}

} // namespace

int lrt_range_rays_impl(lrt_ctx* ctx, int H, int W, const float* inc_table, float inc_lo, float inc_hi, float pixel_offset,
                        float angle_offset, const float* sensor2world, const float* range_map, float* out, float* centre, cudaStream_t s)
{
    if (H <= 0 || W <= 0 || (long long)H * W > 0x7fffffffLL / 4 || !sensor2world || !out) { ctx->set_error("lrt_range_rays: bad size or null argument"); return LRT_ERR_INVALID; }
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RayGrid g;
    g.H = H; g.W = W; g.inc_table = inc_table; g.inc_lo = inc_lo; g.inc_hi = inc_hi; g.pixel_offset = pixel_offset; g.angle_offset = angle_offset; g.s2w = sensor2world;
    const int n = H * W;
    ctx->span_begin("k_range_rays", s);
    if (range_map) k_range_rays<true><<<(n + 255) / 256, 256, 0, s>>>(g, range_map, out, centre);
    else k_range_rays<false><<<(n + 255) / 256, 256, 0, s>>>(g, nullptr, out, centre);
    ctx->span_end(s);
    ctx->launches += 1;
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}
