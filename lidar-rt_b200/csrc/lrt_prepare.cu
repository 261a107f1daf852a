// lrt_prepare.cu — fused parameter activation + world transform + per-asset concatenation, forward and backward.
//
// the step immediately above the tracer. The reference does it with ~20 small torch kernels and
// several P x 232 B copies per call (lib/gaussian_renderer/__init__.py:76-134 over the accessors of
// lib/scene/gaussian_model.py:112-148):
//   means3D   = cat_a( xyz_a @ R_a^T + T_a )                       get_world_xyz, R_a = build_rotation(pose quaternion)
//   opacity   = cat_a( sigmoid(opacity_a) )                        get_opacity
//   scales    = cat_a( exp(scaling_a) )                            get_scaling
//   rotations = normalize(rotation_0)                              background / static scene (:117-118), or
//               cat( normalize(rotation_0), q_a (x) normalize(normalize(rotation_a)) )   dynamic scene (:119-130)
//   shs       = cat_a( cat(features_dc_a, features_rest_a, dim=1) )   get_features
// Here: two launches forward (per-Gaussian parameters; SH rows as a flat coalesced copy) and two backward, reading
// the leaf tensors in place and writing the tracer's inputs / the leaf gradients directly. No gradient flows to
// the actor poses (plain tensors in the reference's BoundingBox.frame).
#include "lrt_ctx.cuh"

namespace {

struct PrepAsset {
    const float *xyz, *scaling, *rotation, *opacity, *dc, *rest, *pose_T, *pose_q;
    float *d_xyz, *d_scaling, *d_rotation, *d_opacity, *d_dc, *d_rest;
    int first, P, compose, pad;
};
struct PrepTable { int n, total, M, pad; PrepAsset a[LRT_MAX_ASSETS]; };

__device__ __forceinline__ int find_asset(const PrepTable& t, int i)
{
    int lo = 0, hi = t.n - 1;                            // last asset whose first index is <= i
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (t.a[mid].first <= i) lo = mid; else hi = mid - 1; }
    return lo;
}

// build_rotation (general_utils.py:176-197) of the pose quaternion; rows of R
__device__ __forceinline__ void pose_matrix(const float* q_, float* R)
{
    const float n = sqrtf(q_[0] * q_[0] + q_[1] * q_[1] + q_[2] * q_[2] + q_[3] * q_[3]);
    const float r = q_[0] / n, x = q_[1] / n, y = q_[2] / n, z = q_[3] / n;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z); R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y); R[7] = 2.f * (y * z + r * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

// torch.nn.functional.normalize(v, dim=1): v / max(|v|, 1e-12)
__device__ __forceinline__ float normalize4(const float* v, float* o)
{
    const float n = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]), 1e-12f);
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = v[k] / n;
    return n;
}
// VJP of normalize4 (inside the clamp's active region the gradient of |v| applies; below 1e-12 it is a plain scale)
__device__ __forceinline__ void normalize4_vjp(const float* o, float n, const float* g, float* dv)
{
    if (n <= 1e-12f) {
#pragma unroll
        for (int k = 0; k < 4; k++) dv[k] = g[k] / n;
        return;
    }
    const float dot = o[0] * g[0] + o[1] * g[1] + o[2] * g[2] + o[3] * g[3];
#pragma unroll
    for (int k = 0; k < 4; k++) dv[k] = (g[k] - o[k] * dot) / n;
}

// Hamilton product, real part first (general_utils.py:156-174)
__device__ __forceinline__ void quat_mul(const float* a, const float* b, float* o)
{
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

__global__ void __launch_bounds__(256) k_prepare(const __grid_constant__ PrepTable t, float* __restrict__ means, float* __restrict__ scales,
                                                 float* __restrict__ rots, float* __restrict__ opac)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.total) return;
    const PrepAsset& a = t.a[find_asset(t, i)];
    const int j = i - a.first;
    const float x[3] = {a.xyz[3 * (size_t)j], a.xyz[3 * (size_t)j + 1], a.xyz[3 * (size_t)j + 2]};
    float m[3] = {x[0], x[1], x[2]};
    if (a.pose_q) {
        const float q[4] = {a.pose_q[0], a.pose_q[1], a.pose_q[2], a.pose_q[3]};
        float R[9];
        pose_matrix(q, R);
#pragma unroll
        for (int k = 0; k < 3; k++) m[k] = (x[0] * R[3 * k] + x[1] * R[3 * k + 1] + x[2] * R[3 * k + 2]) + a.pose_T[k];
    }
    means[3 * (size_t)i] = m[0]; means[3 * (size_t)i + 1] = m[1]; means[3 * (size_t)i + 2] = m[2];
    const float2 s = *reinterpret_cast<const float2*>(a.scaling + 2 * (size_t)j);
    *reinterpret_cast<float2*>(scales + 2 * (size_t)i) = make_float2(expf(s.x), expf(s.y));
    opac[i] = 1.0f / (1.0f + expf(-a.opacity[j]));
    const float4 r4 = *reinterpret_cast<const float4*>(a.rotation + 4 * (size_t)j);
    const float raw[4] = {r4.x, r4.y, r4.z, r4.w};
    float n1[4], out[4];
    normalize4(raw, n1);
    if (a.compose) {
        float n2[4];
        normalize4(n1, n2);                               // the reference normalises the actors' local rotation twice (:128)
        const float q[4] = {a.pose_q ? a.pose_q[0] : 0.f, a.pose_q ? a.pose_q[1] : 0.f, a.pose_q ? a.pose_q[2] : 0.f, a.pose_q ? a.pose_q[3] : 0.f};
        quat_mul(q, n2, out);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) out[k] = n1[k];
    }
    *reinterpret_cast<float4*>(rots + 4 * (size_t)i) = make_float4(out[0], out[1], out[2], out[3]);
}

// SH rows: shs[i][0] = dc[j][0], shs[i][1..M-1] = rest[j][0..M-2]; one thread per float, flat and coalesced.
// BACKWARD = the same mapping the other way (dL_dshs -> d_dc, d_rest).
template <bool BACKWARD>
__global__ void __launch_bounds__(256) k_prepare_sh(const __grid_constant__ PrepTable t, float* __restrict__ shs)
{
    const int row = 3 * t.M;
    const size_t n = (size_t)t.total * row;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / row), c = (int)(e - (size_t)i * row);
        const PrepAsset& a = t.a[find_asset(t, i)];
        const size_t j = (size_t)(i - a.first);
        if (!BACKWARD) shs[e] = c < 3 ? a.dc[3 * j + c] : a.rest[(size_t)(row - 3) * j + (c - 3)];
        else if (c < 3) { if (a.d_dc) a.d_dc[3 * j + c] = shs[e]; }
        else if (a.d_rest) a.d_rest[(size_t)(row - 3) * j + (c - 3)] = shs[e];
    }
}

// The same for rows that are a whole number of 16-byte groups (3 M % 4 == 0, shs 16-byte aligned; M = 16 always in the
// reference): one thread per float4 of the concatenated row, 32-bit index arithmetic, one asset search per four floats.
// The one-float-per-thread form above spends ~80 instructions per float on a 64-bit division and the search and runs at a
// third of the copy bandwidth.
template <bool BACKWARD>
__global__ void __launch_bounds__(256) k_prepare_sh4(const __grid_constant__ PrepTable t, float4* __restrict__ shs4)
{
    const unsigned row = 3u * (unsigned)t.M, row4 = row >> 2;
    const unsigned n4 = (unsigned)t.total * row4;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += gridDim.x * blockDim.x) {
        const unsigned i = e / row4, c = (e - i * row4) << 2;          // Gaussian, first float of this group within its row
        const PrepAsset& a = t.a[find_asset(t, (int)i)];
        const size_t j = (size_t)((int)i - a.first);
        const float* __restrict__ rest = a.rest + (size_t)(row - 3) * j;
        if (!BACKWARD) {
            float4 v;
            if (c == 0) v = make_float4(a.dc[3 * j], a.dc[3 * j + 1], a.dc[3 * j + 2], rest[0]);
            else v = make_float4(rest[c - 3], rest[c - 2], rest[c - 1], rest[c]);
            shs4[e] = v;
        } else {
            const float4 v = shs4[e];
            float* d_rest = a.d_rest ? a.d_rest + (size_t)(row - 3) * j : nullptr;
            if (c == 0) {
                if (a.d_dc) { a.d_dc[3 * j] = v.x; a.d_dc[3 * j + 1] = v.y; a.d_dc[3 * j + 2] = v.z; }
                if (d_rest) d_rest[0] = v.w;
            } else if (d_rest) {
                d_rest[c - 3] = v.x; d_rest[c - 2] = v.y; d_rest[c - 1] = v.z; d_rest[c] = v.w;
            }
        }
    }
}

template <bool BACKWARD>
void launch_prepare_sh(const PrepTable& t, float* shs, cudaStream_t s)
{
    const bool fast = ((3 * t.M) & 3) == 0 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0 && (long long)t.total * (3 * t.M / 4) < 0xffffffffLL - 148 * 32 * 256;
    if (fast) k_prepare_sh4<BACKWARD><<<148 * 32, 256, 0, s>>>(t, reinterpret_cast<float4*>(shs));
    else k_prepare_sh<BACKWARD><<<148 * 16, 256, 0, s>>>(t, shs);
}

__global__ void __launch_bounds__(256) k_prepare_backward(const __grid_constant__ PrepTable t, const float* __restrict__ g_means,
                                                          const float* __restrict__ g_scales, const float* __restrict__ g_rots,
                                                          const float* __restrict__ g_opac)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.total) return;
    const PrepAsset& a = t.a[find_asset(t, i)];
    const size_t j = (size_t)(i - a.first);
    if (a.d_xyz) {
        const float g[3] = {g_means[3 * (size_t)i], g_means[3 * (size_t)i + 1], g_means[3 * (size_t)i + 2]};
        float d[3] = {g[0], g[1], g[2]};
        if (a.pose_q) {                                   // means = R x + T  ->  dx = R^T g
            const float q[4] = {a.pose_q[0], a.pose_q[1], a.pose_q[2], a.pose_q[3]};
            float R[9];
            pose_matrix(q, R);
#pragma unroll
            for (int k = 0; k < 3; k++) d[k] = g[0] * R[k] + g[1] * R[3 + k] + g[2] * R[6 + k];
        }
        a.d_xyz[3 * j] = d[0]; a.d_xyz[3 * j + 1] = d[1]; a.d_xyz[3 * j + 2] = d[2];
    }
    if (a.d_scaling) {                                    // scales = exp(s)
        const float2 s = *reinterpret_cast<const float2*>(a.scaling + 2 * j);
        const float2 g = *reinterpret_cast<const float2*>(g_scales + 2 * (size_t)i);
        *reinterpret_cast<float2*>(a.d_scaling + 2 * j) = make_float2(g.x * expf(s.x), g.y * expf(s.y));
    }
    if (a.d_opacity) {                                    // opacity = sigmoid(x)
        const float o = 1.0f / (1.0f + expf(-a.opacity[j]));
        a.d_opacity[j] = g_opac[i] * (o * (1.0f - o));
    }
    if (a.d_rotation) {
        const float4 r4 = *reinterpret_cast<const float4*>(a.rotation + 4 * j);
        const float4 g4 = *reinterpret_cast<const float4*>(g_rots + 4 * (size_t)i);
        const float raw[4] = {r4.x, r4.y, r4.z, r4.w};
        float g[4] = {g4.x, g4.y, g4.z, g4.w};
        float n1[4], d1[4];
        const float l1 = normalize4(raw, n1);
        if (a.compose) {
            float n2[4];
            const float l2 = normalize4(n1, n2);
            // r = q (x) n2  ->  d_n2 = conj(q) (x) g   (the transpose of left multiplication by q)
            const float qc[4] = {a.pose_q ? a.pose_q[0] : 0.f, a.pose_q ? -a.pose_q[1] : 0.f, a.pose_q ? -a.pose_q[2] : 0.f, a.pose_q ? -a.pose_q[3] : 0.f};
            float dn2[4], dn1[4];
            quat_mul(qc, g, dn2);
            normalize4_vjp(n2, l2, dn2, dn1);
#pragma unroll
            for (int k = 0; k < 4; k++) g[k] = dn1[k];
        }
        normalize4_vjp(n1, l1, g, d1);
        *reinterpret_cast<float4*>(a.d_rotation + 4 * j) = make_float4(d1[0], d1[1], d1[2], d1[3]);
    }
}

int fill_table(lrt_ctx* ctx, int n_assets, const lrt_asset* assets, int M, PrepTable& t, const char* who)
{
    if (n_assets <= 0 || n_assets > LRT_MAX_ASSETS || !assets) { ctx->set_error((std::string(who) + ": need 1..LRT_MAX_ASSETS assets").c_str()); return LRT_ERR_INVALID; }
    if (M < 1 || M > 16) { ctx->set_error((std::string(who) + ": need 1 <= M <= 16").c_str()); return LRT_ERR_INVALID; }
    long long total = 0;
    for (int k = 0; k < n_assets; k++) {
        const lrt_asset& s = assets[k];
        if (s.P <= 0 || !s.xyz || !s.scaling || !s.rotation || !s.opacity || !s.features_dc || (M > 1 && !s.features_rest)) {
            ctx->set_error((std::string(who) + ": asset with P <= 0 or a null leaf tensor").c_str()); return LRT_ERR_INVALID;
        }
        if ((s.pose_T == nullptr) != (s.pose_quat == nullptr)) { ctx->set_error((std::string(who) + ": pose_T and pose_quat go together").c_str()); return LRT_ERR_INVALID; }
        if (s.compose_rotation && !s.pose_quat) { ctx->set_error((std::string(who) + ": compose_rotation needs a pose").c_str()); return LRT_ERR_INVALID; }
        if ((reinterpret_cast<uintptr_t>(s.scaling) & 7) || (reinterpret_cast<uintptr_t>(s.rotation) & 15)) {
            ctx->set_error((std::string(who) + ": scaling must be 8-byte and rotation 16-byte aligned").c_str()); return LRT_ERR_INVALID;
        }
        PrepAsset& a = t.a[k];
        a.xyz = s.xyz; a.scaling = s.scaling; a.rotation = s.rotation; a.opacity = s.opacity; a.dc = s.features_dc; a.rest = s.features_rest;
        a.pose_T = s.pose_T; a.pose_q = s.pose_quat;
        a.d_xyz = s.d_xyz; a.d_scaling = s.d_scaling; a.d_rotation = s.d_rotation; a.d_opacity = s.d_opacity; a.d_dc = s.d_features_dc; a.d_rest = s.d_features_rest;
        a.first = (int)total; a.P = s.P; a.compose = s.compose_rotation ? 1 : 0; a.pad = 0;
        total += s.P;
        if (total > 0x7fffffffLL / 64) { ctx->set_error((std::string(who) + ": too many Gaussians").c_str()); return LRT_ERR_INVALID; }
    }
    t.n = n_assets; t.total = (int)total; t.M = M; t.pad = 0;
    return LRT_OK;
}

} // namespace

int lrt_prepare_impl(lrt_ctx* ctx, int n_assets, const lrt_asset* assets, int M, float* means, float* scales, float* rots,
                     float* opac, float* shs, cudaStream_t s)
{
    if (!means || !scales || !rots || !opac) { ctx->set_error("lrt_prepare: null output"); return LRT_ERR_INVALID; }      // shs == NULL: no concatenated copy (lrt_set_sh_parts)
    if ((reinterpret_cast<uintptr_t>(scales) & 7) || (reinterpret_cast<uintptr_t>(rots) & 15)) { ctx->set_error("lrt_prepare: scales must be 8-byte and rots 16-byte aligned"); return LRT_ERR_INVALID; }
    PrepTable t;
    const int rc = fill_table(ctx, n_assets, assets, M, t, "lrt_prepare");
    if (rc != LRT_OK) return rc;
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->span_begin("k_prepare", s);
    k_prepare<<<(t.total + 255) / 256, 256, 0, s>>>(t, means, scales, rots, opac);
    if (shs) launch_prepare_sh<false>(t, shs, s);
    ctx->span_end(s);
    ctx->launches += 2;
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}

int lrt_prepare_backward_impl(lrt_ctx* ctx, int n_assets, const lrt_asset* assets, int M, const float* g_means, const float* g_scales,
                              const float* g_rots, const float* g_opac, const float* g_shs, cudaStream_t s)
{
    if (!g_means || !g_scales || !g_rots || !g_opac) { ctx->set_error("lrt_prepare_backward: null gradient input"); return LRT_ERR_INVALID; }      // g_shs == NULL: SH gradients were written in place
    if ((reinterpret_cast<uintptr_t>(g_scales) & 7) || (reinterpret_cast<uintptr_t>(g_rots) & 15)) { ctx->set_error("lrt_prepare_backward: dL_dscales must be 8-byte and dL_drots 16-byte aligned"); return LRT_ERR_INVALID; }
    PrepTable t;
    const int rc = fill_table(ctx, n_assets, assets, M, t, "lrt_prepare_backward");
    if (rc != LRT_OK) return rc;
    for (int k = 0; k < n_assets; k++) {
        if ((t.a[k].d_scaling && (reinterpret_cast<uintptr_t>(t.a[k].d_scaling) & 7)) || (t.a[k].d_rotation && (reinterpret_cast<uintptr_t>(t.a[k].d_rotation) & 15))) {
            ctx->set_error("lrt_prepare_backward: d_scaling must be 8-byte and d_rotation 16-byte aligned"); return LRT_ERR_INVALID;
        }
    }
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->span_begin("k_prepare_backward", s);
    k_prepare_backward<<<(t.total + 255) / 256, 256, 0, s>>>(t, g_means, g_scales, g_rots, g_opac);
    if (g_shs) launch_prepare_sh<true>(t, const_cast<float*>(g_shs), s);
    ctx->span_end(s);
    ctx->launches += 2;
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}
