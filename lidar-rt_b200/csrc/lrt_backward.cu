// lrt_backward.cu — backward device program: gradients onto xyz / scale / rotation / opacity / SH.
//
// Replaces __raygen__ot of the reference's backward pipeline
// (submodules/diff-lidar-tracer/optix_tracer/backward.cu:434-691) and TraceSurfelsBackwardCUDA
// (trace_surfels.cpp:268-386). The per-hit math follows backward.cu:577-675,
// compute_transmat_uv_backward (:339-431), computeColorFromSHBackward (:123-247) and
// quat_to_rotmat_vjp (auxiliary.h:389-433) term for term.
//
// Two ways to enumerate a ray's contributing hits, in the same front-to-back order:
//   k_backward_list  : replay the (id, depth) list the forward pass recorded — no traversal at all;
//   k_backward_trace : re-trace through the LBVH exactly like the reference does (used for rays
//                      whose list overflowed `cap`, or when the caller passes no lists).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "lrt_ctx.cuh"
#include "lrt_trace.cuh"

namespace {

struct RayState {
    float g_rgb[3], g_d, g_n[3];          // upstream gradients (dL_daccum / dL_dT are ignored, backward.cu:476)
    float F_c[3], F_d, F_n[3], F_T;       // saved forward outputs
    float C[3], N[3], Dp, T;              // running prefix (includes the current hit)
};

struct GradOut {
    float* d_means; float* d_shs; float* d_opac; float* d_scales; float* d_rots;
    int vec;                              // 128-bit / 64-bit vector reductions usable (alignment checked on the host)
    const ShTab* sh_tab;                  // non-null: SH gradients go straight into the leaf gradients d_features_dc / d_features_rest
    int sh_vec;                           // ... whose `rest` bases are 16-byte aligned (vector reductions on the aligned middle of a row)
};

// sm_90+ vector float reductions: one L2 operation for 4 (2) adjacent floats instead of 4 (2).
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ void cross3(const float* a, const float* b, float* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ __forceinline__ void load_sh_bw(const float* __restrict__ shs, int g, int M, int nb, float* sh)
{
    const float* p = shs + (size_t)g * M * 3;
    if ((M & 3) == 0 && ((reinterpret_cast<uintptr_t>(shs) & 15) == 0)) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const int n4 = (nb * 3 + 3) >> 2;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            if (i < n4) { const float4 v = ld_f4(p4 + i); sh[4 * i] = v.x; sh[4 * i + 1] = v.y; sh[4 * i + 2] = v.z; sh[4 * i + 3] = v.w; }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 48; i++) if (i < nb * 3) sh[i] = ld_f(p + i);
    }
}

// Geometry of one proxy hit of ray (o, d) at depth dpt on Gaussian g (forward.cu:116-141, :228-251 recomputed
// from the raw parameters, same arithmetic as the forward's record).
struct HitGeo { Derived s; float rr[3], u, v, cosv, power, G, alpha; };

__device__ __forceinline__ void hit_geo(int g, float dpt, const float* o, const float* d,
                                        const float* __restrict__ means, const float* __restrict__ scales,
                                        const float* __restrict__ rots, const float* __restrict__ opac, float mod, HitGeo& h)
{
    const float mu_[3] = {ld_f(means + 3 * (size_t)g), ld_f(means + 3 * (size_t)g + 1), ld_f(means + 3 * (size_t)g + 2)};
    const float2 sc2 = __ldg(reinterpret_cast<const float2*>(scales + 2 * (size_t)g));
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(rots + 4 * (size_t)g));
    const float sc_[2] = {sc2.x, sc2.y};
    const float q_[4] = {q4.x, q4.y, q4.z, q4.w};
    derive_surfel(mu_, sc_, q_, ld_f(opac + g), mod, h.s);
    const float xyz[3] = {o[0] + dpt * d[0], o[1] + dpt * d[1], o[2] + dpt * d[2]};
#pragma unroll
    for (int k = 0; k < 3; k++) h.rr[k] = xyz[k] - h.s.mu[k];
    h.u = h.s.Lu[0] * h.rr[0] + h.s.Lu[1] * h.rr[1] + h.s.Lu[2] * h.rr[2];
    h.v = h.s.Lv[0] * h.rr[0] + h.s.Lv[1] * h.rr[1] + h.s.Lv[2] * h.rr[2];
    h.cosv = -((h.s.mu[0] - o[0]) * h.s.n[0] + (h.s.mu[1] - o[1]) * h.s.n[1] + (h.s.mu[2] - o[2]) * h.s.n[2]);
    const float rho = h.u * h.u + h.v * h.v;
    h.power = -0.5f * rho;
    h.G = expf(h.power);
    h.alpha = fminf(LRT_ALPHA_MAX, h.s.op * h.G);
}

// What a hit takes from (and adds to) its ray's running state: blending weight w = alpha T, the prefix sums
// including this hit, and dL/dalpha (backward.cu:577-604). Advances st.T.
__device__ __forceinline__ float hit_state(RayState& st, float alpha, const float* c, const float* n, float dpt,
                                           const float* bg, int flags, float& w_out)
{
    const float T = st.T;
    const float w = alpha * T;
#pragma unroll
    for (int k = 0; k < 3; k++) { st.C[k] += w * c[k]; st.N[k] += w * n[k]; }        // :577-578
    st.Dp += w * dpt;
    const float inv1a = 1.0f / (1.0f - alpha);
    float dalpha = 0.0f;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) dalpha += st.g_rgb[ch] * (T * c[ch] - (st.F_c[ch] - st.C[ch]) * inv1a);      // :590
    if (!(flags & LRT_FLAG_FIX_BG_GRAD)) {
        float dbg = 0.0f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) dbg += st.g_rgb[ch] * bg[ch];
        dalpha += dbg * (-st.F_T * inv1a);                                            // :595-598 (duplicate term)
    }
    dalpha += st.g_d * (T * dpt - (st.F_d - st.Dp) * inv1a);                          // :601
    {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) acc += st.g_n[k] * (T * n[k] - (st.F_n[k] - st.N[k]) * inv1a);
        dalpha += acc;                                                                // :604
    }
    st.T = T * (1.0f - alpha);
    w_out = w;
    return dalpha;
}

// Everything downstream of dL/dalpha for one hit: VJPs onto opacity, scale, rotation, mean (incl. the gradient
// through the hit depth) and SH, scattered with float reductions (backward.cu:607-675).
struct HitGrads { float d_opac, dmu[3], dsc[2], dq[4]; };

__device__ __forceinline__ void hit_vjp(const float* o, const float* d, const HitGeo& h, float w, float dalpha,
                                        float g_d, const float* g_n, HitGrads& out)
{
    const Derived& s = h.s;
    const float G = h.G, u = h.u, v = h.v;
    const float* rr = h.rr;
    const float dD_gs = g_d * w;
    float dN_gs[3];
#pragma unroll
    for (int k = 0; k < 3; k++) dN_gs[k] = g_n[k] * w;
    if (s.op * G > LRT_ALPHA_MAX) dalpha = 0.0f;                                      // :607-608
    const float dG = s.op * dalpha;
    out.d_opac = G * dalpha;                                                          // :615
    const float nsign = h.cosv > 0.0f ? 1.0f : -1.0f;                                 // :649-650

    // compute_transmat_uv_backward (:339-431)
    const float du = dG * -G * u, dv = dG * -G * v;
    float dtu[3], dtv[3], dn[3], dmu[3], dsc[2], dxyz[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { dtu[k] = du * rr[k] / s.sx; dtv[k] = dv * rr[k] / s.sy; dn[k] = dN_gs[k] * nsign; }
    dsc[0] = dG * (G * u * u / s.sx); dsc[1] = dG * (G * v * v / s.sy);
#pragma unroll
    for (int k = 0; k < 3; k++) { dmu[k] = dG * (G * (s.Lu[k] * u + s.Lv[k] * v)); dxyz[k] = du * s.Lu[k] + dv * s.Lv[k]; }
    const float dd = (dxyz[0] * d[0] + dxyz[1] * d[1] + dxyz[2] * d[2]) + dD_gs;      // :390
    // gradient through the hit depth: plane of the proxy triangle that contains the hit
    // (:628-647: even -> corners 0,1,2 ; odd -> corners 1,2,3 ; corners = build2DRectangle order)
    const bool odd = !(v >= u);
    const float ax = s.sx * s.f, ay = s.sy * s.f;
    const float lx[4] = {-1.f, -1.f, 1.f, 1.f}, ly[4] = {1.f, -1.f, 1.f, -1.f};
    const int i1 = odd ? 1 : 0, i2 = odd ? 2 : 1, i3 = odd ? 3 : 2;
    float v1[3], v2[3], v3[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = s.tu[k] * ax, b = s.tv[k] * ay;
        v1[k] = (lx[i1] * a + ly[i1] * b) + s.mu[k];
        v2[k] = (lx[i2] * a + ly[i2] * b) + s.mu[k];
        v3[k] = (lx[i3] * a + ly[i3] * b) + s.mu[k];
    }
    const float cutoff = (float)(sqrt(2.0 * log((double)s.op * 255.)) + 0.01);       // :625 (double there too)
    const float h1x = lx[i1] * cutoff, h1y = ly[i1] * cutoff, h2x = lx[i2] * cutoff, h2y = ly[i2] * cutoff,
                h3x = lx[i3] * cutoff, h3y = ly[i3] * cutoff;
    float e21[3], e31[3], e23[3], e12[3], nT[3], cT[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        e21[k] = v2[k] - v1[k]; e31[k] = v3[k] - v1[k]; e23[k] = v2[k] - v3[k]; e12[k] = v1[k] - v2[k];
        cT[k] = v1[k] - o[k];
    }
    cross3(e21, e31, nT);
    const float pp = nT[0] * cT[0] + nT[1] * cT[1] + nT[2] * cT[2];
    const float qq = nT[0] * d[0] + nT[1] * d[1] + nT[2] * d[2];
    float a_[3], x1[3], x2[3], x3[3];
#pragma unroll
    for (int k = 0; k < 3; k++) a_[k] = (cT[k] - pp / qq * d[k]) / qq;                // :399
    cross3(e23, a_, x1); cross3(e31, a_, x2); cross3(e12, a_, x3);
    float Sx[3], Sy[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float dv1 = x1[k] * dd + nT[k] / qq * dd, dv2 = x2[k] * dd, dv3 = x3[k] * dd;   // :400-402
        Sx[k] = h1x * dv1 + h2x * dv2 + h3x * dv3;
        Sy[k] = h1y * dv1 + h2y * dv2 + h3y * dv3;
        dtu[k] += s.sx * Sx[k]; dtv[k] += s.sy * Sy[k];                               // :405-410
        dmu[k] += dv1 + dv2 + dv3;                                                    // :430
    }
    dsc[0] += (s.sx * s.Lu[0]) * Sx[0] + (s.sx * s.Lu[1]) * Sx[1] + (s.sx * s.Lu[2]) * Sx[2];   // :426-427
    dsc[1] += (s.sy * s.Lv[0]) * Sy[0] + (s.sy * s.Lv[1]) * Sy[1] + (s.sy * s.Lv[2]) * Sy[2];

    // quat_to_rotmat_vjp (auxiliary.h:389-433); Gm[r][c] = dL/dR(r,c), columns (tu, tv, n)
    {
        const float qw = s.qn[0], qx = s.qn[1], qy = s.qn[2], qz = s.qn[3];
#define GM(r, c) ((c) == 0 ? dtu[r] : (c) == 1 ? dtv[r] : dn[r])
        const float dq0 = 2.0f * (qx * (GM(2, 1) - GM(1, 2)) + qy * (GM(0, 2) - GM(2, 0)) + qz * (GM(1, 0) - GM(0, 1)));
        const float dq1 = 2.0f * (-2.0f * qx * (GM(1, 1) + GM(2, 2)) + qy * (GM(1, 0) + GM(0, 1)) + qz * (GM(2, 0) + GM(0, 2)) + qw * (GM(2, 1) - GM(1, 2)));
        const float dq2 = 2.0f * (qx * (GM(1, 0) + GM(0, 1)) - 2.0f * qy * (GM(0, 0) + GM(2, 2)) + qz * (GM(2, 1) + GM(1, 2)) + qw * (GM(0, 2) - GM(2, 0)));
        const float dq3 = 2.0f * (qx * (GM(2, 0) + GM(0, 2)) + qy * (GM(2, 1) + GM(1, 2)) - 2.0f * qz * (GM(0, 0) + GM(1, 1)) + qw * (GM(1, 0) - GM(0, 1)));
#undef GM
        out.dq[0] = dq0; out.dq[1] = dq1; out.dq[2] = dq2; out.dq[3] = dq3;
    }
    out.dsc[0] = dsc[0]; out.dsc[1] = dsc[1];
    out.dmu[0] = dmu[0]; out.dmu[1] = dmu[1]; out.dmu[2] = dmu[2];
}

// scatter one hit's gradients with float reductions (backward.cu:659-675)
__device__ __forceinline__ void hit_scatter(int g, const float* o, const float* d, const HitGeo& h, float w, float dalpha,
                                            float g_d, const float* g_n, float* dcol, bool clamped0, const float* basis, int nb, int M,
                                            const GradOut& go)
{
    HitGrads hv;
    hit_vjp(o, d, h, w, dalpha, g_d, g_n, hv);
    atomicAdd(go.d_opac + g, hv.d_opac);
    if (go.vec) {
        red_add_v2(go.d_scales + 2 * (size_t)g, hv.dsc[0], hv.dsc[1]);
        red_add_v4(go.d_rots + 4 * (size_t)g, hv.dq[0], hv.dq[1], hv.dq[2], hv.dq[3]);
    } else {
        atomicAdd(go.d_scales + 2 * (size_t)g, hv.dsc[0]); atomicAdd(go.d_scales + 2 * (size_t)g + 1, hv.dsc[1]);
        atomicAdd(go.d_rots + 4 * (size_t)g, hv.dq[0]); atomicAdd(go.d_rots + 4 * (size_t)g + 1, hv.dq[1]);
        atomicAdd(go.d_rots + 4 * (size_t)g + 2, hv.dq[2]); atomicAdd(go.d_rots + 4 * (size_t)g + 3, hv.dq[3]);
    }
    atomicAdd(go.d_means + 3 * (size_t)g, hv.dmu[0]); atomicAdd(go.d_means + 3 * (size_t)g + 1, hv.dmu[1]);
    atomicAdd(go.d_means + 3 * (size_t)g + 2, hv.dmu[2]);
    // computeColorFromSHBackward (:123-247): dL_dsh[j] = basis_j * dL_dcolour, channel 0 zero if clamped
    if (clamped0) dcol[0] = 0.0f;
    if (go.sh_tab) {                                   // in-place leaf gradients (rare paths: scalar reductions)
        int j;
        const ShPartDev& p = sh_find(go.sh_tab, g, j);
#pragma unroll
        for (int jj = 0; jj < 16; jj++) {
            if (jj < nb) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    float* dst = jj == 0 ? (p.d_dc ? p.d_dc + 3 * (size_t)j + ch : nullptr)
                                         : (p.d_rest ? p.d_rest + 3 * (size_t)(go.sh_tab->M - 1) * j + 3 * (jj - 1) + ch : nullptr);
                    if (dst) atomicAdd(dst, basis[jj] * dcol[ch]);
                }
            }
        }
        return;
    }
    float* dsh = go.d_shs + (size_t)g * M * 3;
    if (go.vec && (M & 3) == 0) {
        // (j, ch) pairs are contiguous: 3 nb floats = groups of 4 (nb = 1 -> 3 floats: scalar tail)
        const int nf = 3 * nb;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            float vals[4];
#pragma unroll
            for (int e = 0; e < 4; e++) { const int idx = 4 * i + e; vals[e] = basis[idx / 3] * dcol[idx % 3]; }
            if (4 * i + 3 < nf) red_add_v4(dsh + 4 * i, vals[0], vals[1], vals[2], vals[3]);
            else {
#pragma unroll
                for (int e = 0; e < 4; e++) if (4 * i + e < nf) atomicAdd(dsh + 4 * i + e, vals[e]);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (j < nb) {
                atomicAdd(dsh + 3 * j, basis[j] * dcol[0]);
                atomicAdd(dsh + 3 * j + 1, basis[j] * dcol[1]);
                atomicAdd(dsh + 3 * j + 2, basis[j] * dcol[2]);
            }
        }
    }
}

// One proxy hit of ray (o, d) at depth dpt on Gaussian g, start to finish (serial replay / re-trace paths).
// FILTER = true replays the forward's skip / termination rules (re-trace path);
// returns 0 = not contributing, 1 = composited (gradients scattered), 2 = ray terminated.
template <bool FILTER>
__device__ __forceinline__ int hit_backward(int g, float dpt, const float* o, const float* d, const float* dirn,
                                            const float* __restrict__ means, const float* __restrict__ scales,
                                            const float* __restrict__ rots, const float* __restrict__ opac,
                                            const float* __restrict__ shs, int D, int M, float mod,
                                            const float* __restrict__ bg, int flags, RayState& st, const GradOut& go)
{
    if (FILTER && dpt < LRT_MIN_T) return 0;                                          // backward.cu:528
    HitGeo h;
    hit_geo(g, dpt, o, d, means, scales, rots, opac, mod, h);
    if (FILTER && h.power > 0.0f) return 0;
    if (FILTER && h.alpha < 1.0f / 255.0f) return 0;
    if (FILTER && st.T * (1.0f - h.alpha) < LRT_T_MIN) return 2;
    const int nb = (D + 1) * (D + 1);
    float sh[48], c[3], basis[16]; bool clamped0;
    if (go.sh_tab) load_sh_parts(go.sh_tab, g, nb, sh);
    else load_sh_bw(shs, g, M, nb, sh);
    sh_colour<true>(D, dirn, sh, c, clamped0, basis);
    float w;
    const float dalpha = hit_state(st, h.alpha, c, h.s.n, dpt, bg, flags, w);
    float dcol[3] = {st.g_rgb[0] * w, st.g_rgb[1] * w, st.g_rgb[2] * w};
    hit_scatter(g, o, d, h, w, dalpha, st.g_d, st.g_n, dcol, clamped0, basis, nb, M, go);
    return 1;
}

__device__ __forceinline__ void ray_state_init(RayState& st, int r, const float* __restrict__ fwd_out,
                                               const float* __restrict__ dL)
{
    const float* g = dL + (size_t)LRT_NCH * r;
    const float* f = fwd_out + (size_t)LRT_NCH * r;
#pragma unroll
    for (int k = 0; k < 3; k++) { st.g_rgb[k] = g[k]; st.g_n[k] = g[5 + k]; st.F_c[k] = f[k]; st.F_n[k] = f[5 + k]; st.C[k] = 0.f; st.N[k] = 0.f; }
    st.g_d = g[3]; st.F_d = f[3]; st.F_T = f[8]; st.Dp = 0.f; st.T = 1.f;
}

struct BwArgs {
    int R; const float* ray_o; int ray_o_stride; const float* ray_d; const float* bg;
    const float* means; const float* scales; const float* rots; const float* opac; const float* shs;
    int D, M; float mod; const float* fwd_out; const float* dL; int flags;
    const int32_t* hit_gidx; const float* hit_t; const float4* hit_aux; const int32_t* hit_cnt; int cap;
    const int* order;                     // rays by descending hit count, or nullptr
    const int* only_flag;                 // k_backward_list: if set, run only when the device flag is non-zero
    GradOut go;
};

__global__ void __launch_bounds__(128) k_iota(int n, int* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

__global__ void __launch_bounds__(128) k_backward_list(BwArgs a)
{
    const int s_ = blockIdx.x * blockDim.x + threadIdx.x;
    if (s_ >= a.R) return;
    const int r = a.order ? a.order[s_] : s_;
    const int cnt = a.hit_cnt[r];
    if (cnt <= 0 || cnt > a.cap) return;                 // overflowed rays are handled by k_backward_trace
    if (a.only_flag && *a.only_flag == 0) return;
    const float o[3] = {a.ray_o[(size_t)r * a.ray_o_stride], a.ray_o[(size_t)r * a.ray_o_stride + 1], a.ray_o[(size_t)r * a.ray_o_stride + 2]};
    const float d[3] = {a.ray_d[3 * (size_t)r], a.ray_d[3 * (size_t)r + 1], a.ray_d[3 * (size_t)r + 2]};
    const float dl = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float dirn[3] = {d[0] / dl, d[1] / dl, d[2] / dl};
    const float bg[3] = {a.bg[0], a.bg[1], a.bg[2]};
    RayState st;
    ray_state_init(st, r, a.fwd_out, a.dL);
    for (int k = 0; k < cnt; k++) {
        const int g = a.hit_gidx[(size_t)k * a.R + r];
        const float dpt = a.hit_t[(size_t)k * a.R + r];
        hit_backward<false>(g, dpt, o, d, dirn, a.means, a.scales, a.rots, a.opac, a.shs, a.D, a.M, a.mod, bg, a.flags, st, a.go);
    }
}

// ---- grouped replay (default when the forward recorded hit_aux) -------------------------------------------------
// The serial part of a ray's backward is tiny: transmittance and the running sums of w c, w depth (w n) in front
// of each hit. With (alpha, colour) of every contributing hit recorded by the forward, that part needs no Gaussian
// parameter at all. And the expensive part — VJPs and the scatter of 58 floats per hit — is bound by the L2's
// reduction rate (18 x 16-byte red operations per hit, ~100 G/s measured), while every Gaussian is hit by several
// neighbouring rays. So the hits are regrouped BY GAUSSIAN before they are scattered:
//   k_bw_gcount + scan : hits per Gaussian -> first flat position of every Gaussian's group
//   k_bw_prefix        : one thread per ray walks its recorded (depth, alpha, colour) and drops, per hit, the record
//                        (g, ray, depth, w, dL/dalpha) into the Gaussian's group. ~30 flops and 24 B per hit.
//   k_bw_hits          : one thread per record, all independent: recompute the surfel frame, VJPs; lanes holding the
//                        same Gaussian are neighbours, so a segmented warp reduction leaves ONE set of reductions per
//                        (Gaussian, warp). Parameter loads coalesce for the same reason, and the SH coefficients are
//                        never read (their gradient only needs the basis and dL/dcolour).
struct BwFlat { int2* a; float4* b; int* gcnt; const int* goff; int P; int capacity; int* legacy_flag; };

__global__ void __launch_bounds__(128) k_bw_gcount(BwArgs a, int* __restrict__ gcnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const int cnt = a.hit_cnt[r];
    if (cnt <= 0 || cnt > a.cap) return;                  // overflowed rays are handled by k_backward_trace
    for (int k = 0; k < cnt; k++) atomicAdd(gcnt + a.hit_gidx[(size_t)k * a.R + r], 1);
}

__global__ void __launch_bounds__(128) k_bw_prefix(BwArgs a, BwFlat f)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    if (f.goff[f.P] > f.capacity) { if (r == 0) *f.legacy_flag = 1; return; }      // records do not fit: serial replay of everything
    const int cnt = a.hit_cnt[r];
    if (cnt <= 0 || cnt > a.cap) return;
    const float bg[3] = {a.bg[0], a.bg[1], a.bg[2]};
    RayState st;
    ray_state_init(st, r, a.fwd_out, a.dL);
    const bool need_n = st.g_n[0] != 0.0f || st.g_n[1] != 0.0f || st.g_n[2] != 0.0f;    // dL/dnormal is zero in practice (train.py never feeds it)
    int g_nx = a.hit_gidx[r]; float t_nx = a.hit_t[r]; float4 ax_nx = a.hit_aux[r];
    for (int k = 0; k < cnt; k++) {
        const int g = g_nx;
        const float dpt = t_nx;
        const float4 ax = ax_nx;
        if (k + 1 < cnt) {                                 // next hit's record: in flight while this one is folded
            const size_t at = (size_t)(k + 1) * a.R + r;
            g_nx = a.hit_gidx[at]; t_nx = a.hit_t[at]; ax_nx = a.hit_aux[at];
        }
        const int pos = f.goff[g] + atomicSub(f.gcnt + g, 1) - 1;          // a free slot of g's group (order within a group is irrelevant)
        const float c[3] = {ax.y, ax.z, ax.w};
        float n[3] = {0.f, 0.f, 0.f};
        if (need_n) {                                      // the normal prefix only matters then; same arithmetic as derive_surfel
            const float4 q4 = __ldg(reinterpret_cast<const float4*>(a.rots + 4 * (size_t)g));
            const float nrm = q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w;
            const float inv = 1.0f / sqrtf(nrm);
            const float w_ = q4.x * inv, x = q4.y * inv, y = q4.z * inv, z = q4.w * inv;
            n[0] = 2.0f * (x * z + w_ * y); n[1] = 2.0f * (y * z - w_ * x); n[2] = 1.0f - 2.0f * (x * x + y * y);
        }
        float w;
        const float dalpha = hit_state(st, ax.x, c, n, dpt, bg, a.flags, w);
        const unsigned clamp = __float_as_uint(ax.y) & 0x80000000u;        // c0 == -0.0: channel 0 was clamped
        f.a[pos] = make_int2(g, (int)((unsigned)r | clamp));
        f.b[pos] = make_float4(dpt, w, dalpha, 0.0f);
    }
}

// sum of v over the lanes to the right that hold the same Gaussian (equal ids are contiguous): lane i ends up with the
// total of [i, end of its segment]; `same` bit s = lane + 2^s is in range and holds the same id
__device__ __forceinline__ float seg_sum(float v, unsigned same)
{
#pragma unroll
    for (int s = 0; s < 5; s++) { const float o = __shfl_down_sync(0xffffffffu, v, 1 << s); if (same & (1u << s)) v += o; }
    return v;
}

#define BW_SROW 49
template <bool INPLACE>                                  // SH gradients into the concatenated (P, M, 3) buffer, or in place into the leaf gradients
__global__ void __launch_bounds__(256, 3) k_bw_hits(BwArgs a, BwFlat f)      // 3 blocks per SM (<= 85 registers) measured faster than 2 at 103
{
    extern __shared__ float s_row[];                     // in-place SH gradients only: 8 warps x 32 lanes x BW_SROW floats
    const unsigned FULL = 0xffffffffu;
    const int n = f.goff[f.P];
    if (n > f.capacity) return;
    const int nb = (a.D + 1) * (a.D + 1);
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n; base += gridDim.x * blockDim.x) {
        const int h = base + lane;
        int g = -1 - lane;                                 // inactive lanes: distinct ids, never equal to a neighbour
        HitGrads hv; hv.d_opac = 0.f; hv.dmu[0] = hv.dmu[1] = hv.dmu[2] = 0.f; hv.dsc[0] = hv.dsc[1] = 0.f; hv.dq[0] = hv.dq[1] = hv.dq[2] = hv.dq[3] = 0.f;
        float dcol[3] = {0.f, 0.f, 0.f}, basis[16];
#pragma unroll
        for (int j = 0; j < 16; j++) basis[j] = 0.0f;
        if (h < n) {
            const int2 ra = f.a[h];
            const float4 rb = f.b[h];
            g = ra.x;
            const int r = ra.y & 0x7fffffff;
            const float o[3] = {a.ray_o[(size_t)r * a.ray_o_stride], a.ray_o[(size_t)r * a.ray_o_stride + 1], a.ray_o[(size_t)r * a.ray_o_stride + 2]};
            const float d[3] = {a.ray_d[3 * (size_t)r], a.ray_d[3 * (size_t)r + 1], a.ray_d[3 * (size_t)r + 2]};
            const float dl = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            const float dirn[3] = {d[0] / dl, d[1] / dl, d[2] / dl};
            const float* gl = a.dL + (size_t)LRT_NCH * r;
            const float g_n[3] = {gl[5], gl[6], gl[7]};
            const float w = rb.y;
            dcol[0] = (ra.y < 0) ? 0.0f : gl[0] * w; dcol[1] = gl[1] * w; dcol[2] = gl[2] * w;     // channel 0 clamped: no SH gradient (:134)
            HitGeo hg;
            hit_geo(g, rb.x, o, d, a.means, a.scales, a.rots, a.opac, a.mod, hg);
            sh_basis(a.D, dirn, basis);
#pragma unroll
            for (int j = 0; j < 16; j++) if (j >= nb) basis[j] = 0.0f;
            hit_vjp(o, d, hg, w, rb.z, gl[3], g_n, hv);
        }
        // segment structure of this warp's 32 records
        unsigned same = 0;
#pragma unroll
        for (int s = 0; s < 5; s++) { const int og = __shfl_down_sync(FULL, g, 1 << s); if (lane + (1 << s) < 32 && og == g) same |= 1u << s; }
        const int pg = __shfl_up_sync(FULL, g, 1);
        const bool head = (h < n) && (lane == 0 || pg != g);
        const GradOut& go = a.go;
        {
            const float t0 = seg_sum(hv.d_opac, same);
            const float m0 = seg_sum(hv.dmu[0], same), m1 = seg_sum(hv.dmu[1], same), m2 = seg_sum(hv.dmu[2], same);
            const float s0 = seg_sum(hv.dsc[0], same), s1 = seg_sum(hv.dsc[1], same);
            const float q0 = seg_sum(hv.dq[0], same), q1 = seg_sum(hv.dq[1], same), q2 = seg_sum(hv.dq[2], same), q3 = seg_sum(hv.dq[3], same);
            if (head) {
                atomicAdd(go.d_opac + g, t0);
                atomicAdd(go.d_means + 3 * (size_t)g, m0); atomicAdd(go.d_means + 3 * (size_t)g + 1, m1); atomicAdd(go.d_means + 3 * (size_t)g + 2, m2);
                if (go.vec) { red_add_v2(go.d_scales + 2 * (size_t)g, s0, s1); red_add_v4(go.d_rots + 4 * (size_t)g, q0, q1, q2, q3); }
                else {
                    atomicAdd(go.d_scales + 2 * (size_t)g, s0); atomicAdd(go.d_scales + 2 * (size_t)g + 1, s1);
                    atomicAdd(go.d_rots + 4 * (size_t)g, q0); atomicAdd(go.d_rots + 4 * (size_t)g + 1, q1);
                    atomicAdd(go.d_rots + 4 * (size_t)g + 2, q2); atomicAdd(go.d_rots + 4 * (size_t)g + 3, q3);
                }
            }
        }
        // SH: dL_dsh[j][ch] = basis_j dL_dcolour[ch], four contiguous floats at a time
        const int nf = 3 * nb;
        if (INPLACE) {
            // leaf gradients in place: a Gaussian's row is d_features_dc[j] (3 floats) + d_features_rest[j] (3 (M - 1) floats at a
            // 4-byte aligned address). The segment heads park their reduced row in shared memory (the 16-byte groups of the target
            // do not line up with the groups the reduction produces), then emit it: red.v4 over the aligned middle of the rest
            // part, scalar reductions for the up to three floats in front of and behind it and for the dc part — at most 20
            // reductions per row against 12 for an aligned concatenated row.
            float* srow = s_row + ((threadIdx.x >> 5) * 32 + lane) * BW_SROW;      // odd stride: the lanes' rows start in different banks
#pragma unroll
            for (int i = 0; i < 12; i++) {
                if (4 * i < nf) {                          // warp-uniform
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int idx = 4 * i + e;
                        const float v = seg_sum(basis[idx / 3] * dcol[idx % 3], same);
                        if (head) srow[idx] = v;
                    }
                }
            }
            __syncwarp(FULL);
            if (head) {                                    // every head emits its own row; the alignment only decides where the 16-byte groups start
                const int rf = 3 * (go.sh_tab->M - 1), nrest = nf - 3;
                int j;
                const ShPartDev& p = sh_find(go.sh_tab, g, j);
                const size_t f0 = (size_t)rf * j;
                float* rest = p.d_rest ? p.d_rest + f0 : nullptr;
                const int lead = go.sh_vec ? min((4 - (int)(f0 & 3)) & 3, nrest) : nrest;      // floats in front of the first aligned group
                const int nv4 = go.sh_vec ? (nrest - lead) >> 2 : 0;
                const int tail = nrest - lead - 4 * nv4;
                // straight-line, predicated: heads of different alignment stay in step (loops with alignment-dependent trip counts
                // made every alignment class run the emission on its own, 2-3 lanes at a time: ncu, profiles/r2_o_*)
                if (p.d_dc) {
                    float* dc = p.d_dc + 3 * (size_t)j;
                    atomicAdd(dc, srow[0]); atomicAdd(dc + 1, srow[1]); atomicAdd(dc + 2, srow[2]);
                }
                if (rest) {
                    const float* q = srow + 3 + lead;
                    float* r4 = rest + lead;
#pragma unroll
                    for (int k = 0; k < 3; k++) if (k < lead) atomicAdd(rest + k, srow[3 + k]);
#pragma unroll
                    for (int v = 0; v < 11; v++) if (v < nv4) red_add_v4(r4 + 4 * v, q[4 * v], q[4 * v + 1], q[4 * v + 2], q[4 * v + 3]);
#pragma unroll
                    for (int k = 0; k < 3; k++) if (k < tail) atomicAdd(r4 + 4 * nv4 + k, q[4 * nv4 + k]);
                }
            }
            __syncwarp(FULL);
            continue;
        }
        float* dsh = go.d_shs + (size_t)(g < 0 ? 0 : g) * a.M * 3;
        const bool vec = go.vec && (a.M & 3) == 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            if (4 * i < nf) {                              // warp-uniform
                float vals[4];
#pragma unroll
                for (int e = 0; e < 4; e++) { const int idx = 4 * i + e; vals[e] = seg_sum(basis[idx / 3] * dcol[idx % 3], same); }
                if (head) {
                    if (vec && 4 * i + 3 < nf) red_add_v4(dsh + 4 * i, vals[0], vals[1], vals[2], vals[3]);
                    else {
#pragma unroll
                        for (int e = 0; e < 4; e++) if (4 * i + e < nf) atomicAdd(dsh + 4 * i + e, vals[e]);
                    }
                }
            }
        }
    }
}

// One WARP per ray: the ray's recorded hits are processed 32 at a time, one hit per lane. What a hit needs
// from its predecessors — transmittance T_k and the running sums of w c, w n, w depth — comes from warp
// scans (a product scan of (1 - alpha) and seven sum scans), after which every lane runs the same per-hit
// VJP code as the serial replay. The scans associate differently from the serial loop (ulp-level), which
// is below the float-atomic reordering noise the gradients carry anyway.
__global__ void __launch_bounds__(128) k_backward_warp(BwArgs a)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < a.R; r += warps_per_grid) {
        const int cnt = a.hit_cnt[r];
        if (cnt <= 0 || cnt > a.cap) continue;              // overflowed rays are handled by k_backward_trace
        const float o[3] = {a.ray_o[(size_t)r * a.ray_o_stride], a.ray_o[(size_t)r * a.ray_o_stride + 1], a.ray_o[(size_t)r * a.ray_o_stride + 2]};
        const float d[3] = {a.ray_d[3 * (size_t)r], a.ray_d[3 * (size_t)r + 1], a.ray_d[3 * (size_t)r + 2]};
        const float dl = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const float dirn[3] = {d[0] / dl, d[1] / dl, d[2] / dl};
        const float bg[3] = {a.bg[0], a.bg[1], a.bg[2]};
        RayState base;
        ray_state_init(base, r, a.fwd_out, a.dL);
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int k = k0 + lane;
            const bool active = k < cnt;
            int g = 0; float dpt = 0.f;
            float alpha = 0.f, wc[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f};
            if (active) {
                g = a.hit_gidx[(size_t)k * a.R + r];
                dpt = a.hit_t[(size_t)k * a.R + r];
                // pass 1: alpha, colour and normal of this hit (same arithmetic as hit_backward)
                const float mu_[3] = {ld_f(a.means + 3 * (size_t)g), ld_f(a.means + 3 * (size_t)g + 1), ld_f(a.means + 3 * (size_t)g + 2)};
                const float2 sc2 = __ldg(reinterpret_cast<const float2*>(a.scales + 2 * (size_t)g));
                const float4 q4 = __ldg(reinterpret_cast<const float4*>(a.rots + 4 * (size_t)g));
                const float sc_[2] = {sc2.x, sc2.y};
                const float q_[4] = {q4.x, q4.y, q4.z, q4.w};
                Derived s;
                derive_surfel(mu_, sc_, q_, ld_f(a.opac + g), a.mod, s);
                const float rr0 = (o[0] + dpt * d[0]) - s.mu[0], rr1 = (o[1] + dpt * d[1]) - s.mu[1], rr2 = (o[2] + dpt * d[2]) - s.mu[2];
                const float u = s.Lu[0] * rr0 + s.Lu[1] * rr1 + s.Lu[2] * rr2;
                const float v = s.Lv[0] * rr0 + s.Lv[1] * rr1 + s.Lv[2] * rr2;
                alpha = fminf(LRT_ALPHA_MAX, s.op * expf(-0.5f * (u * u + v * v)));
                if (!a.go.sh_tab && (a.M & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.shs) & 15) == 0)) {
                    sh_colour_stream(a.D, dirn, a.shs + (size_t)g * a.M * 3, wc);
                } else {
                    float sh[48]; bool cl;
                    if (a.go.sh_tab) load_sh_parts(a.go.sh_tab, g, (a.D + 1) * (a.D + 1), sh);
                    else load_sh_bw(a.shs, g, a.M, (a.D + 1) * (a.D + 1), sh);
                    sh_colour<false>(a.D, dirn, sh, wc, cl, nullptr);
                }
                nrm[0] = s.n[0]; nrm[1] = s.n[1]; nrm[2] = s.n[2];
            }
            // transmittance before each hit: exclusive product scan of (1 - alpha)
            float om = active ? 1.0f - alpha : 1.0f, pr = om;
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) { const float t = __shfl_up_sync(FULL, pr, sft); if (lane >= sft) pr *= t; }
            float excl = __shfl_up_sync(FULL, pr, 1); if (lane == 0) excl = 1.0f;
            const float T = base.T * excl;
            const float w = alpha * T;
            // exclusive sums of w c, w n, w depth
            float v7[7] = {w * wc[0], w * wc[1], w * wc[2], w * nrm[0], w * nrm[1], w * nrm[2], w * dpt};
            float ex7[7], tot7[7];
#pragma unroll
            for (int j = 0; j < 7; j++) {
                float x = v7[j];
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) { const float t = __shfl_up_sync(FULL, x, sft); if (lane >= sft) x += t; }
                tot7[j] = __shfl_sync(FULL, x, 31);
                ex7[j] = x - v7[j];
            }
            if (active) {
                RayState st = base;
                st.T = T;
                st.C[0] = base.C[0] + ex7[0]; st.C[1] = base.C[1] + ex7[1]; st.C[2] = base.C[2] + ex7[2];
                st.N[0] = base.N[0] + ex7[3]; st.N[1] = base.N[1] + ex7[4]; st.N[2] = base.N[2] + ex7[5];
                st.Dp = base.Dp + ex7[6];
                hit_backward<false>(g, dpt, o, d, dirn, a.means, a.scales, a.rots, a.opac, a.shs, a.D, a.M, a.mod, bg, a.flags, st, a.go);
            }
            // carry to the next 32 hits
            base.T *= __shfl_sync(FULL, pr, 31);
            base.C[0] += tot7[0]; base.C[1] += tot7[1]; base.C[2] += tot7[2];
            base.N[0] += tot7[3]; base.N[1] += tot7[4]; base.N[2] += tot7[5];
            base.Dp += tot7[6];
        }
    }
}

// only_overflow: process just the rays whose forward list overflowed
__global__ void __launch_bounds__(128) k_backward_trace(BvhView bvh, BwArgs a, int only_overflow)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    if (only_overflow && a.hit_cnt[r] <= a.cap) return;
    const float o[3] = {a.ray_o[(size_t)r * a.ray_o_stride], a.ray_o[(size_t)r * a.ray_o_stride + 1], a.ray_o[(size_t)r * a.ray_o_stride + 2]};
    const float d[3] = {a.ray_d[3 * (size_t)r], a.ray_d[3 * (size_t)r + 1], a.ray_d[3 * (size_t)r + 2]};
    const float dl = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float dirn[3] = {d[0] / dl, d[1] / dl, d[2] / dl};
    const float bg[3] = {a.bg[0], a.bg[1], a.bg[2]};
    RayState st;
    ray_state_init(st, r, a.fwd_out, a.dL);
    float base = 0.f, dpt = 0.f;
    int last = -1;
    for (;;) {
        RaySetup rs;
        ray_setup(rs, o, d, base);
        unsigned long long kb[LRT_KBUF];
#ifdef LRT_STATS
        int nv_ = 0;
        const int n = trace_round(bvh, rs, kb, nv_);
#else
        const int n = trace_round(bvh, rs, kb);
#endif
        unsigned long long hits[LRT_KBUF];
#pragma unroll
        for (int i = 0; i < LRT_KBUF; i++) hits[i] = kb[i];
        bool terminated = false;
        for (int i = 0; i < n; i++) {
            const unsigned long long key = hits[i];
            const int g = (int)(unsigned)(key & 0xffffffffull);
            dpt = __uint_as_float((unsigned)(key >> 32)) + base;
            if (dpt < LRT_MIN_T) continue;                                            // backward.cu:528
            if (g == last) continue;                                                  // :534-536
            last = g;
            const int res = hit_backward<true>(g, dpt, o, d, dirn, a.means, a.scales, a.rots, a.opac, a.shs, a.D, a.M, a.mod,
                                               bg, a.flags, st, a.go);
            if (res == 2) { terminated = true; break; }
        }
        if (terminated || n < LRT_KBUF) break;
        base = (float)((double)dpt + LRT_STEP_EPS);
    }
}

} // namespace

int lrt_backward_impl(lrt_ctx* ctx, int R, const float* ray_o, int ray_o_stride, const float* ray_d,
                      const float* bg, int P, const float* means, const float* scales, const float* rots,
                      const float* opac, const float* shs, int D, int M, float mod,
                      const float* fwd_out, const float* dL_dout,
                      const int32_t* hit_gidx, const float* hit_t, const float* hit_aux, const int32_t* hit_cnt, int cap,
                      float* dL_dmeans, float* dL_dshs, float* dL_dopac, float* dL_dscales,
                      float* dL_drots, int flags, cudaStream_t s)
{
    const ShTab* sh_tab = nullptr;
    if (!shs || !dL_dshs) {                              // SH rows and their gradients in place (lrt_set_sh_parts)
        if (shs || dL_dshs) { ctx->set_error("lrt_backward: shs and dL_dshs must both be NULL to use the bound SH parts"); return LRT_ERR_INVALID; }
        if (!ctx->sh_parts_n || ctx->sh_parts_P != P || ctx->sh_parts_M != M) { ctx->set_error("lrt_backward: shs is NULL and no matching SH parts are bound (lrt_set_sh_parts)"); return LRT_ERR_STATE; }
        sh_tab = (const ShTab*)ctx->sh_tab.p;
    }
    if (R < 0 || P <= 0 || !ray_o || !ray_d || !bg || !means || !scales || !rots || !opac || !fwd_out || !dL_dout ||
        !dL_dmeans || !dL_dopac || !dL_dscales || !dL_drots) { ctx->set_error("lrt_backward: null argument"); return LRT_ERR_INVALID; }
    if (ray_o_stride != 0 && ray_o_stride != 3) { ctx->set_error("lrt_backward: ray_o_stride must be 0 or 3"); return LRT_ERR_INVALID; }
    if (D < 0 || D > 3 || M < (D + 1) * (D + 1)) { ctx->set_error("lrt_backward: need 0 <= D <= 3 and M >= (D+1)^2"); return LRT_ERR_INVALID; }
    const bool have_lists = hit_gidx && hit_t && hit_cnt && cap > 0;
    if (!have_lists && (hit_gidx || hit_t)) { ctx->set_error("lrt_backward: hit_gidx, hit_t, hit_cnt and cap go together"); return LRT_ERR_INVALID; }
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // The gradient buffers are zero-filled (464 MB at 2 M Gaussians: pure bandwidth). In the grouped replay nothing touches them
    // before k_bw_hits, so the fill runs on a side stream beside the count / scan / prefix passes, which wait on latency and
    // leave the memory system idle; fork and join are events, so the same shape is recorded when the caller captures a graph.
    const bool grouped = R > 0 && hit_gidx && hit_t && hit_cnt && cap > 0 && ctx->opt_backward_kernel == 2 && hit_aux &&
                         (reinterpret_cast<uintptr_t>(hit_aux) & 15) == 0;
    cudaStream_t zs = s;
    if (grouped) {
        if (!ctx->side_stream) {
            LRT_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
            LRT_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            LRT_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        }
        zs = ctx->side_stream;
        LRT_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, s));
        LRT_CUDA_TRY(ctx, cudaStreamWaitEvent(zs, ctx->ev_fork, 0));
    }
    LRT_CUDA_TRY(ctx, cudaMemsetAsync(dL_dmeans, 0, sizeof(float) * (size_t)P * 3, zs));
    if (sh_tab) {
        for (int k = 0; k < ctx->sh_parts_n; k++) {
            const lrt_sh_part& pt = ctx->sh_parts[k];
            if (pt.d_features_dc) LRT_CUDA_TRY(ctx, cudaMemsetAsync(pt.d_features_dc, 0, sizeof(float) * 3 * (size_t)pt.P, zs));
            if (pt.d_features_rest && M > 1) LRT_CUDA_TRY(ctx, cudaMemsetAsync(pt.d_features_rest, 0, sizeof(float) * 3 * (size_t)(M - 1) * pt.P, zs));
        }
    } else {
        LRT_CUDA_TRY(ctx, cudaMemsetAsync(dL_dshs, 0, sizeof(float) * (size_t)P * M * 3, zs));
    }
    LRT_CUDA_TRY(ctx, cudaMemsetAsync(dL_dopac, 0, sizeof(float) * (size_t)P, zs));
    LRT_CUDA_TRY(ctx, cudaMemsetAsync(dL_dscales, 0, sizeof(float) * (size_t)P * 2, zs));
    LRT_CUDA_TRY(ctx, cudaMemsetAsync(dL_drots, 0, sizeof(float) * (size_t)P * 4, zs));
    if (grouped) LRT_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, zs));
    if (R == 0) return LRT_OK;
    BwArgs a;
    a.R = R; a.ray_o = ray_o; a.ray_o_stride = ray_o_stride; a.ray_d = ray_d; a.bg = bg;
    a.means = means; a.scales = scales; a.rots = rots; a.opac = opac; a.shs = shs;
    a.D = D; a.M = M; a.mod = mod; a.fwd_out = fwd_out; a.dL = dL_dout; a.flags = flags;
    a.hit_gidx = hit_gidx; a.hit_t = hit_t; a.hit_aux = reinterpret_cast<const float4*>(hit_aux); a.hit_cnt = hit_cnt; a.cap = cap;
    a.order = nullptr; a.only_flag = nullptr;
    a.go.d_means = dL_dmeans; a.go.d_shs = dL_dshs; a.go.d_opac = dL_dopac; a.go.d_scales = dL_dscales; a.go.d_rots = dL_drots;
    a.go.vec = ctx->opt_vector_atomics && (sh_tab || (reinterpret_cast<uintptr_t>(dL_dshs) & 15) == 0) &&
               (reinterpret_cast<uintptr_t>(dL_drots) & 15) == 0 && (reinterpret_cast<uintptr_t>(dL_dscales) & 7) == 0;
    a.go.sh_tab = sh_tab; a.go.sh_vec = sh_tab && ctx->opt_vector_atomics && ctx->sh_parts_grad_vec;
    const int TB = 128, GB = (R + TB - 1) / TB;
    const bool can_trace = ctx->built && ctx->P == P && ctx->scale_modifier == mod;
    if (have_lists) {
        if (ctx->num_sms == 0) {
            int sms = 0;
            LRT_CUDA_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
            ctx->num_sms = sms > 0 ? sms : 148;
        }
        if (grouped) {
            // grouped replay: hits regrouped by Gaussian, per-ray prefix pass, one thread per hit + segmented reduction
            long long want = (long long)R * 64; if (want < (1LL << 20)) want = 1LL << 20;
            const long long worst = (long long)R * cap;
            if (want > worst) want = worst;
            const int capacity = (int)(want > 0x7fffff00LL ? 0x7fffff00LL : want);
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bw_off, sizeof(int) * 2 * ((size_t)P + 1) + 64));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bw_rec_a, sizeof(int2) * (size_t)capacity));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bw_rec_b, sizeof(float4) * (size_t)capacity));
            int* gcnt = (int*)ctx->bw_off.p; int* goff = gcnt + (P + 1); int* legacy_flag = goff + (P + 1);
            size_t tb = 0;
            LRT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int*)gcnt, goff, P + 1, s));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bw_sort_tmp, tb));
            LRT_CUDA_TRY(ctx, cudaMemsetAsync(gcnt, 0, sizeof(int) * ((size_t)P + 1), s));
            LRT_CUDA_TRY(ctx, cudaMemsetAsync(legacy_flag, 0, sizeof(int), s));
            BwFlat f;
            f.a = (int2*)ctx->bw_rec_a.p; f.b = (float4*)ctx->bw_rec_b.p; f.gcnt = gcnt; f.goff = goff; f.P = P; f.capacity = capacity; f.legacy_flag = legacy_flag;
            ctx->span_begin("k_bw_gcount", s);
            k_bw_gcount<<<GB, TB, 0, s>>>(a, gcnt);
            LRT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->bw_sort_tmp.p, tb, (const int*)gcnt, goff, P + 1, s));
            ctx->span_end(s);
            ctx->span_begin("k_bw_prefix", s); k_bw_prefix<<<GB, TB, 0, s>>>(a, f); ctx->span_end(s);
            LRT_CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->ev_join, 0));      // the zero-filled gradient buffers
            ctx->span_begin("k_bw_hits", s);
            if (sh_tab) {
                LRT_CUDA_TRY(ctx, cudaFuncSetAttribute(k_bw_hits<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * 8 * 32 * BW_SROW)));
                k_bw_hits<true><<<ctx->num_sms * 8, 256, sizeof(float) * 8 * 32 * BW_SROW, s>>>(a, f);
            } else {
                k_bw_hits<false><<<ctx->num_sms * 8, 256, 0, s>>>(a, f);
            }
            ctx->span_end(s);
            a.only_flag = legacy_flag;                // set on the device if the records did not fit (normally not)
            ctx->span_begin("k_backward_list", s); k_backward_list<<<GB, TB, 0, s>>>(a); ctx->span_end(s);
            ctx->launches += 6;
        } else if (ctx->opt_backward_kernel == 1) {
            if (ctx->num_sms == 0) {
                int sms = 0;
                LRT_CUDA_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
                ctx->num_sms = sms > 0 ? sms : 148;
            }
            ctx->span_begin("k_backward_warp", s); k_backward_warp<<<min((R + 3) / 4, ctx->num_sms * 16), TB, 0, s>>>(a); ctx->span_end(s);
        } else {
            if (ctx->opt_sort_rays) {                  // rays by descending contributing-hit count: equal loop lengths within a warp
                LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bw_ids, sizeof(int) * (size_t)R * 2));
                LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bw_keys, sizeof(int) * (size_t)R));
                int* ids = (int*)ctx->bw_ids.p; int* order = ids + R;
                k_iota<<<GB, TB, 0, s>>>(R, ids);
                size_t tb = 0;
                LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, (const int*)hit_cnt, (int*)ctx->bw_keys.p, (const int*)ids, order, R, 0, 12, s));
                LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bw_sort_tmp, tb));
                ctx->span_begin("ray_order_sort", s);
                LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairsDescending(ctx->bw_sort_tmp.p, tb, (const int*)hit_cnt, (int*)ctx->bw_keys.p, (const int*)ids, order, R, 0, 12, s));
                ctx->span_end(s);
                a.order = order;
            }
            ctx->span_begin("k_backward_list", s); k_backward_list<<<GB, TB, 0, s>>>(a); ctx->span_end(s);
        }
        ctx->launches += 1;
        // Rays whose list overflowed need the structure that produced them. The caller (host
        // wrapper) guarantees it is current; without one they cannot be differentiated.
        if (can_trace) { ctx->span_begin("k_backward_trace", s); k_backward_trace<<<GB, TB, 0, s>>>(ctx->view(), a, 1); ctx->span_end(s); ctx->launches += 1; }
    } else {
        if (!can_trace) { ctx->set_error("lrt_backward: no hit lists given and no matching acceleration structure built"); return LRT_ERR_STATE; }
        ctx->span_begin("k_backward_trace", s); k_backward_trace<<<GB, TB, 0, s>>>(ctx->view(), a, 0); ctx->span_end(s);
        ctx->launches += 1;
    }
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}
