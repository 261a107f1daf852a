// lrt_common.cuh — shared device/host definitions of the B200-native LiDAR surfel tracer.
//
// Arithmetic contract: this library is compiled with -fmad=false, so every expression below is
// evaluated exactly as written (IEEE fp32, no implicit FMA contraction). The expressions that
// decide WHICH surfels a ray hits and in WHAT order (derive_surfel, quad test, depth) are written
// in the same operation order as the parity oracle, so hit lists are reproducible bit for bit.
// Explicit fmaf() is used only where results are not part of that contract (box slabs).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define LRT_NCH 9                 // reference: optix_tracer/config.h:24 NUM_CHANNELS_F
#define LRT_KBUF 16               // reference: optix_tracer/config.h:16 CHUNK_SIZE
#define LRT_STEP_EPS 0.00001      // reference: optix_tracer/config.h:17 STEP_EPSILON (double literal)
#define LRT_TMAX 1e16f            // reference: forward.cu:55
#define LRT_MIN_T 0.2f            // reference: forward.cu:214
#define LRT_ALPHA_MAX 0.99f       // reference: forward.cu:249
#define LRT_T_MIN 0.0001f         // reference: forward.cu:254
#define LRT_MAX_LEVELS 8          // 8^8 = 16.7 M surfels; one trail byte per level fits 64 bits
#define LRT_WIDTH 8               // children per node

// SH constants, reference: optix_tracer/auxiliary.h:23-40
#define LRT_SH_C0 0.28209479177387814f
#define LRT_SH_C1 0.4886025119029199f
#define LRT_SH_C2_0 1.0925484305920792f
#define LRT_SH_C2_1 -1.0925484305920792f
#define LRT_SH_C2_2 0.31539156525252005f
#define LRT_SH_C2_3 -1.0925484305920792f
#define LRT_SH_C2_4 0.5462742152960396f
#define LRT_SH_C3_0 -0.5900435899266435f
#define LRT_SH_C3_1 2.890611442640554f
#define LRT_SH_C3_2 -0.4570457994644658f
#define LRT_SH_C3_3 0.3731763325901154f
#define LRT_SH_C3_4 -0.4570457994644658f
#define LRT_SH_C3_5 1.445305721320277f
#define LRT_SH_C3_6 -0.5900435899266435f

// One surfel in Morton order: 64 B = 4 x 128-bit loads.
//   r0 = (mu.x, mu.y, mu.z, f)        f = proxy cutoff sqrt(2 ln(255 o)) + 0.01
//   r1 = (Lu.x, Lu.y, Lu.z, opacity)  Lu = tu / sx  (row 0 of S^-1 R^T, forward.cu:130-132)
//   r2 = (Lv.x, Lv.y, Lv.z, gidx)     Lv = tv / sy ; gidx = caller's index (int bits)
//   r3 = (n.x,  n.y,  n.z,  unused)   n = R[:,2]
struct __align__(16) SurfelRec { float4 r0, r1, r2, r3; };

// 8-wide node: child boxes as SoA, 192 B = 12 x 128-bit loads.
struct __align__(16) Node8 { float lox[8], loy[8], loz[8], hix[8], hiy[8], hiz[8]; };

// Compact leaf for the wavefront's leaf kernel: the 8 surfel boxes of a level-0 node quantised to 8 bits per
// coordinate inside the leaf's own box (80 B instead of 192 B; the kernel is L2-bandwidth bound). Conservative:
// decoded lo <= true lo and decoded hi >= true hi, verified at build time with the decode expression itself.
struct __align__(16) LeafQ {
    float lo[3], sc[3];            // box = lo + q * sc
    unsigned char qlo[3][8], qhi[3][8];
    float pad[2];
};

struct BvhView {
    const SurfelRec* rec;         // (P_pad)
    const Node8* nodes;           // all levels, level 0 (children = surfels) first
    const SurfelRec* rec_g;       // (P) the same records indexed by the caller's Gaussian id (compositing gathers by id: no
                                  // id -> Morton position lookup in front of every record fetch)
    const LeafQ* leafq;           // (P_pad / 8) compact level-0 nodes
    int level_off[LRT_MAX_LEVELS];
    int levels;
    int P;
};

struct Derived {
    float mu[3], tu[3], tv[3], n[3], Lu[3], Lv[3];
    float sx, sy, op, f;
    float qn[4];
};

// Proxy + surfel frame of one Gaussian. Same operation order as oracle derive().
// reference: general_utils.py:176-197 (build_rotation), auxiliary.h:306-328, :445-452,
//            primitive_utils.py:182-224 (cutoff), forward.cu:116-141 (L = S^-1 R^T)
__device__ __forceinline__ void derive_surfel(const float* mu, const float* sc, const float* q, float op,
                                              float mod, Derived& o)
{
    const float nrm = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const float inv = 1.0f / sqrtf(nrm);          // the reference's rsqrtf is approximate; IEEE here
    const float w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    o.qn[0] = w; o.qn[1] = x; o.qn[2] = y; o.qn[3] = z;
    o.tu[0] = 1.0f - 2.0f * (y * y + z * z); o.tv[0] = 2.0f * (x * y - w * z); o.n[0] = 2.0f * (x * z + w * y);
    o.tu[1] = 2.0f * (x * y + w * z); o.tv[1] = 1.0f - 2.0f * (x * x + z * z); o.n[1] = 2.0f * (y * z - w * x);
    o.tu[2] = 2.0f * (x * z - w * y); o.tv[2] = 2.0f * (y * z + w * x); o.n[2] = 1.0f - 2.0f * (x * x + y * y);
    o.sx = sc[0]; o.sy = sc[1]; o.op = op;
    const float isx = 1.0f / (mod * sc[0]), isy = 1.0f / (mod * sc[1]);
#pragma unroll
    for (int k = 0; k < 3; k++) { o.mu[k] = mu[k]; o.Lu[k] = o.tu[k] * isx; o.Lv[k] = o.tv[k] * isy; }
    o.f = sqrtf(2.0f * logf(op * 255.0f)) + 0.01f;
}

// SH colour of the (normalised) ray direction + the basis values (for the VJP).
// reference: forward.cu:67-111 / backward.cu:68-118. sh = (M,3) coefficients of one Gaussian.
template <bool WITH_BASIS>
__device__ __forceinline__ void sh_colour(int deg, const float* dirn, const float* __restrict__ sh,
                                          float* c, bool& clamped0, float* basis)
{
    const float x = dirn[0], y = dirn[1], z = dirn[2];
    float b[16];
    int nb = 1;
    b[0] = LRT_SH_C0;
    if (deg > 0) {
        b[1] = -LRT_SH_C1 * y; b[2] = LRT_SH_C1 * z; b[3] = -LRT_SH_C1 * x; nb = 4;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = LRT_SH_C2_0 * xy; b[5] = LRT_SH_C2_1 * yz; b[6] = LRT_SH_C2_2 * (2.0f * zz - xx - yy);
            b[7] = LRT_SH_C2_3 * xz; b[8] = LRT_SH_C2_4 * (xx - yy); nb = 9;
            if (deg > 2) {
                b[9] = LRT_SH_C3_0 * y * (3.0f * xx - yy);
                b[10] = LRT_SH_C3_1 * xy * z;
                b[11] = LRT_SH_C3_2 * y * (4.0f * zz - xx - yy);
                b[12] = LRT_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = LRT_SH_C3_4 * x * (4.0f * zz - xx - yy);
                b[14] = LRT_SH_C3_5 * z * (xx - yy);
                b[15] = LRT_SH_C3_6 * x * (xx - 3.0f * yy);
                nb = 16;
            }
        }
    }
    float r0 = b[0] * sh[0], r1 = b[0] * sh[1], r2 = b[0] * sh[2];
#pragma unroll
    for (int j = 1; j < 16; j++) {
        if (j < nb) { r0 = r0 + b[j] * sh[3 * j]; r1 = r1 + b[j] * sh[3 * j + 1]; r2 = r2 + b[j] * sh[3 * j + 2]; }
    }
    c[0] = r0 + 0.5f; c[1] = r1 + 0.5f; c[2] = r2 + 0.5f;
    clamped0 = c[0] < 0.0f;
    if (clamped0) c[0] = -0.0f;        // clamped to zero; the sign bit carries "was clamped" to the recorded hit list (x + -0 == x)
    if (WITH_BASIS) {
#pragma unroll
        for (int j = 0; j < 16; j++) basis[j] = j < nb ? b[j] : 0.0f;
    }
}

__device__ __forceinline__ float ld_f(const float* p) { return __ldg(p); }

// SH rows read (and differentiated) IN PLACE from the model's leaf tensors instead of a concatenated (P, M, 3) copy
// (lrt_set_sh_parts): asset k owns the Gaussians [first, first + P) of the concatenation; row j of it is
// cat(features_dc[j] (1,3), features_rest[j] (M-1,3)) — gaussian_model.py:141-144. Lives in device memory (ctx->sh_tab).
struct ShPartDev { int first, P; const float* dc; const float* rest; float* d_dc; float* d_rest; };
struct ShTab { int n, M; ShPartDev part[LRT_MAX_ASSETS]; };

__device__ __forceinline__ const ShPartDev& sh_find(const ShTab* __restrict__ t, int g, int& j)
{
    int lo = 0, hi = t->n - 1;                           // last part whose first index is <= g
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (t->part[mid].first <= g) lo = mid; else hi = mid - 1; }
    j = g - t->part[lo].first;
    return t->part[lo];
}

// the first 3 nb floats of Gaussian g's concatenated row, gathered from the two leaf tensors (generic path: scalar loads)
__device__ __forceinline__ void load_sh_parts(const ShTab* __restrict__ t, int g, int nb, float* sh)
{
    int j;
    const ShPartDev& p = sh_find(t, g, j);
    const float* dc = p.dc + 3 * (size_t)j;
    const float* rest = p.rest + 3 * (size_t)(t->M - 1) * j;
    sh[0] = ld_f(dc); sh[1] = ld_f(dc + 1); sh[2] = ld_f(dc + 2);
#pragma unroll
    for (int i = 3; i < 48; i++) if (i < nb * 3) sh[i] = ld_f(rest + (i - 3));
}

// SH basis only (same expressions as sh_colour); returns the number of active coefficients.
__device__ __forceinline__ int sh_basis(int deg, const float* dirn, float* b)
{
    const float x = dirn[0], y = dirn[1], z = dirn[2];
    int nb = 1;
    b[0] = LRT_SH_C0;
    if (deg > 0) {
        b[1] = -LRT_SH_C1 * y; b[2] = LRT_SH_C1 * z; b[3] = -LRT_SH_C1 * x; nb = 4;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = LRT_SH_C2_0 * xy; b[5] = LRT_SH_C2_1 * yz; b[6] = LRT_SH_C2_2 * (2.0f * zz - xx - yy);
            b[7] = LRT_SH_C2_3 * xz; b[8] = LRT_SH_C2_4 * (xx - yy); nb = 9;
            if (deg > 2) {
                b[9] = LRT_SH_C3_0 * y * (3.0f * xx - yy);
                b[10] = LRT_SH_C3_1 * xy * z;
                b[11] = LRT_SH_C3_2 * y * (4.0f * zz - xx - yy);
                b[12] = LRT_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = LRT_SH_C3_4 * x * (4.0f * zz - xx - yy);
                b[14] = LRT_SH_C3_5 * z * (xx - yy);
                b[15] = LRT_SH_C3_6 * x * (xx - 3.0f * yy);
                nb = 16;
            }
        }
    }
    return nb;
}

// SH colour streamed straight from global memory (no 48-float staging array): the (M,3) row of one
// Gaussian is consumed in memory order, which is coefficient order per channel, so every channel sum is
// accumulated in exactly the order sh_colour() uses. Needs a 16-byte aligned row (M % 4 == 0).
__device__ __forceinline__ void sh_colour_stream(int deg, const float* dirn, const float* __restrict__ row, float* c)
{
    float b[16];
    const int nb = sh_basis(deg, dirn, b);
    const int nf = 3 * nb;
    const float4* p4 = reinterpret_cast<const float4*>(row);
    float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 12; i++) {
        if (4 * i < nf) {
            const float4 v = __ldg(p4 + i);
            const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int idx = 4 * i + e, j = idx / 3, ch = idx % 3;
                if (idx < nf) r[ch] = (j == 0) ? b[0] * vals[e] : r[ch] + b[j] * vals[e];
            }
        }
    }
    c[0] = r[0] + 0.5f; c[1] = r[1] + 0.5f; c[2] = r[2] + 0.5f;
    if (c[0] < 0.0f) c[0] = -0.0f;     // see sh_colour
}
// the same with the basis already evaluated (it depends on the ray only): nb basis values in b. All loads of the row are
// issued before the first use, so a thread waits for the row once, not once per group of loads the scheduler happens to form.
__device__ __forceinline__ void sh_colour_stream_b(int nb, const float* b, const float* __restrict__ row, float* c)
{
    const int nf = 3 * nb;
    const float4* p4 = reinterpret_cast<const float4*>(row);
    float4 v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) v[i] = (4 * i < nf) ? __ldg(p4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 12; i++) {
        if (4 * i < nf) {
            const float vals[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int idx = 4 * i + e, j = idx / 3, ch = idx % 3;
                if (idx < nf) r[ch] = (j == 0) ? b[0] * vals[e] : r[ch] + b[j] * vals[e];
            }
        }
    }
    c[0] = r[0] + 0.5f; c[1] = r[1] + 0.5f; c[2] = r[2] + 0.5f;
    if (c[0] < 0.0f) c[0] = -0.0f;     // see sh_colour
}
__device__ __forceinline__ float4 ld_f4(const float4* p) { return __ldg(p); }

#define LRT_CUDA_TRY(ctx, call)                                                                     \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) { (ctx)->set_error(#call, e__); return LRT_ERR_CUDA; }              \
    } while (0)
