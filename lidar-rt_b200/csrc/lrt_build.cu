// lrt_build.cu — LBVH build / refit over the per-Gaussian proxy quads.
//
// Replaces build2DRectangle (lib/utils/primitive_utils.py:182-224: 4 verts + 2 triangles per
// Gaussian, 72 B written and re-read) and optixAccelBuild/optixAccelCompact
// (submodules/diff-lidar-tracer/trace_surfels.cpp:46-148) of the reference.
//
// Pipeline (all on the caller's stream, no host sync):
//   k_bounds   : min/max of the means                          read 12 B/Gaussian
//   k_morton*  : space-filling-curve key of each mean + identity   read 12 B, write 8-12 B
//                (32-bit cubic-cell keys by default; 30-bit and 63-bit Morton selectable)
//   cub radix  : sort (key, index) pairs, 8 bits per pass
//   k_records  : gather raw parameters through the permutation, derive the surfel frame,
//                write the 64 B record in Morton order (and once more by caller id) and the padded
//                quad AABB into its level-0 node slot           read 40 B, write 2 x 64 + 24 B
//   k_fit      : one launch per upper level: child box = union of the child's 8 boxes
// The hierarchy is IMPLICIT: level l node j has children 8j..8j+7 of level l-1 (level 0: surfels),
// so there are no child pointers, a refit is k_records + k_fit with the stored permutation, and
// the traversal needs no stack (see lrt_trace.cuh).
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>
#include "lrt_ctx.cuh"

namespace {

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_bounds_init(int* b)
{
    if (threadIdx.x < 3) b[threadIdx.x] = 0x7fffffff;          // min
    else if (threadIdx.x < 6) b[threadIdx.x] = (int)0x80000000; // max
}

// Scene box. A thread takes FOUR means at a time as three 128-bit loads (the array is read once, fully coalesced); partial boxes
// meet in the warp, then in the block, and one set of six atomics per BLOCK reaches memory (one per warp was 57 k atomics on six
// addresses, which serialise in the L2: 0.045 ms for a 24 MB read).
__global__ void __launch_bounds__(256) k_bounds(int P, const float* __restrict__ means, int* __restrict__ b)
{
    __shared__ float s_lo[8][3], s_hi[8][3];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    const bool vec = (reinterpret_cast<uintptr_t>(means) & 15) == 0;
    const int nquad = vec ? P >> 2 : 0;                         // groups of four means = three float4
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nquad; i += gridDim.x * blockDim.x) {
        const float4* p4 = reinterpret_cast<const float4*>(means) + 3 * (size_t)i;
        const float4 a = __ldg(p4), c = __ldg(p4 + 1), e = __ldg(p4 + 2);
        const float v[12] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w, e.x, e.y, e.z, e.w};
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const float x = v[j];
            if (x == x && fabsf(x) < 1e30f) { lo[j % 3] = fminf(lo[j % 3], x); hi[j % 3] = fmaxf(hi[j % 3], x); }
        }
    }
    for (int i = 4 * nquad + blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = means[3 * (size_t)i + k];
            if (v == v && fabsf(v) < 1e30f) { lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { s_lo[wib][k] = lo[k]; s_hi[wib][k] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int k = threadIdx.x;
        float l = s_lo[0][k], h = s_hi[0][k];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) { l = fminf(l, s_lo[w][k]); h = fmaxf(h, s_hi[w][k]); }
        atomicMin(&b[k], f2ord(l)); atomicMax(&b[3 + k], f2ord(h));
    }
}

__device__ __forceinline__ unsigned expand10(unsigned v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void __launch_bounds__(256) k_morton(int P, const float* __restrict__ means, const int* __restrict__ b,
                                                unsigned* __restrict__ keys, unsigned* __restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    unsigned q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float lo = ord2f(b[k]), hi = ord2f(b[3 + k]);
        const float ext = fmaxf(hi - lo, 1e-20f);
        float t = (means[3 * i + k] - lo) / ext * 1024.0f;
        t = fminf(fmaxf(t, 0.0f), 1023.0f);
        q[k] = (t == t) ? (unsigned)t : 0u;
    }
    keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    idx[i] = (unsigned)i;
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v)
{
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

// 63-bit keys on CUBIC cells (all axes normalised by the largest extent): a street scene is 220 x 50 x 17 m,
// per-axis normalisation would spend as many bits on 17 m of height as on 220 m of road.
__global__ void __launch_bounds__(256) k_morton64(int P, const float* __restrict__ means, const int* __restrict__ b,
                                                  unsigned long long* __restrict__ keys, unsigned* __restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float lo[3] = {ord2f(b[0]), ord2f(b[1]), ord2f(b[2])};
    const float ext = fmaxf(fmaxf(ord2f(b[3]) - lo[0], ord2f(b[4]) - lo[1]), fmaxf(ord2f(b[5]) - lo[2], 1e-20f));
    unsigned long long q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float t = (means[3 * i + k] - lo[k]) / ext * 2097152.0f;
        t = fminf(fmaxf(t, 0.0f), 2097151.0f);
        q[k] = (t == t) ? (unsigned long long)t : 0ull;
    }
    keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    idx[i] = (unsigned)i;
}

// 32-bit keys with the bits dealt to the axes so that cells stay (nearly) cubic: starting from the scene box,
// every key bit halves the currently longest cell edge (ties: x, y, z). A street scene (220 x 50 x 17 m) gets
// 12/10/10-ish bits instead of 10/10/10 on stretched cells, at half the radix-sort passes of the 63-bit keys.
// plan[bit] = (axis << 8) | shift : key bit (31 - bit) is bit `shift` of the quantised coordinate of `axis`.
// Computed once per build by one thread from the scene box; plan[32..34] = bits per axis.
__global__ void k_morton_plan(const int* __restrict__ b, int* __restrict__ plan)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float cell[3]; int nb[3] = {0, 0, 0}; int ax_of[32];
    for (int k = 0; k < 3; k++) cell[k] = fmaxf(ord2f(b[3 + k]) - ord2f(b[k]), 1e-20f);
    for (int bit = 0; bit < 32; bit++) {
        int ax = 0;
        if (cell[1] > cell[ax]) ax = 1;
        if (cell[2] > cell[ax]) ax = 2;
        ax_of[bit] = ax; cell[ax] *= 0.5f; nb[ax]++;
    }
    int used[3] = {0, 0, 0};
    for (int bit = 0; bit < 32; bit++) { const int ax = ax_of[bit]; used[ax]++; plan[bit] = (ax << 8) | (nb[ax] - used[ax]); }
    for (int k = 0; k < 3; k++) plan[32 + k] = nb[k];
}

__global__ void __launch_bounds__(256) k_morton32(int P, const float* __restrict__ means, const int* __restrict__ b,
                                                  const int* __restrict__ plan, unsigned* __restrict__ keys, unsigned* __restrict__ idx, int nbits)
{
    __shared__ int s_plan[35];
    if (threadIdx.x < 35) s_plan[threadIdx.x] = plan[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    unsigned q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float lo = ord2f(b[k]), ext = fmaxf(ord2f(b[3 + k]) - lo, 1e-20f);
        const int nbk = s_plan[32 + k];
        const float scale = (float)(1u << min(nbk, 24));
        float t = (means[3 * i + k] - lo) / ext * scale;
        t = fminf(fmaxf(t, 0.0f), scale - 1.0f);
        q[k] = (t == t) ? (unsigned)t : 0u;
        if (nbk > 24) q[k] <<= (nbk - 24);
    }
    unsigned key = 0;                                // only the `nbits` top bits the radix sort looks at are assembled
#pragma unroll 8
    for (int bit = 0; bit < nbits; bit++) {
        const int pl = s_plan[bit], ax = pl >> 8, sh = pl & 0xff;
        const unsigned v = ax == 0 ? q[0] : (ax == 1 ? q[1] : q[2]);
        key = (key << 1) | ((v >> sh) & 1u);
    }
    keys[i] = nbits < 32 ? key << (32 - nbits) : key;
    idx[i] = (unsigned)i;
}

__device__ __forceinline__ void rec_store(SurfelRec* dst, const SurfelRec& r)
{
    float4* d = reinterpret_cast<float4*>(dst);
    __stcs(d, r.r0); __stcs(d + 1, r.r1); __stcs(d + 2, r.r2); __stcs(d + 3, r.r3);
}

__global__ void __launch_bounds__(256) k_records(int P, int P_pad, const unsigned* __restrict__ perm,
                                                 const float* __restrict__ means, const float* __restrict__ scales,
                                                 const float* __restrict__ rots, const float* __restrict__ opac,
                                                 float mod, SurfelRec* __restrict__ rec, Node8* __restrict__ leaf,
                                                 SurfelRec* __restrict__ rec_g, LeafQ* __restrict__ leafq,
                                                 Node8* __restrict__ lvl1, int n1, Node8* __restrict__ lvl2, int n2)
{
    // No thread leaves early: the grid covers every child slot of levels 1 and 2 (slots beyond the last leaf are written as
    // empty), the shuffles below run on full warps and the block meets at a barrier.
    __shared__ float s_box[8][6];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // empty slot marker: lo = hi = +FLT_MAX on every axis. A slab test maps it to an interval beyond
    // any tmax (or before 0), so it is never entered; k_fit skips it in unions.
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    if (i < P) {
        const int g = (int)perm[i];
        const float mu[3] = {means[3 * g], means[3 * g + 1], means[3 * g + 2]};
        const float2 sc2 = *reinterpret_cast<const float2*>(scales + 2 * g);
        const float4 q4 = *reinterpret_cast<const float4*>(rots + 4 * g);
        const float sc[2] = {sc2.x, sc2.y};
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
        const float op = opac[g];
        Derived d;
        derive_surfel(mu, sc, q, op, mod, d);
        SurfelRec r;
        r.r0 = make_float4(d.mu[0], d.mu[1], d.mu[2], d.f);
        r.r1 = make_float4(d.Lu[0], d.Lu[1], d.Lu[2], d.op);
        r.r2 = make_float4(d.Lv[0], d.Lv[1], d.Lv[2], __int_as_float(g));
        r.r3 = make_float4(d.n[0], d.n[1], d.n[2], 0.0f);
        // streaming stores: 256 MB of records pass through the L2 once; the 80 MB of raw parameters, gathered through the
        // permutation one 32-byte sector at a time (4-5 sectors per Gaussian), should be what stays resident
        rec_store(rec + i, r);
        rec_store(rec_g + g, r);
        const bool valid = (d.f == d.f) && d.f >= 0.0f && d.f < 1e30f;
        if (valid) {
            const float ax = mod * d.sx * d.f, ay = mod * d.sy * d.f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float e = fabsf(d.tu[k]) * ax + fabsf(d.tv[k]) * ay;
                const float pad = 1e-4f + 1e-5f * (fabsf(mu[k]) + e);       // slab tests are a filter, the quad test decides
                lo[k] = mu[k] - e - pad; hi[k] = mu[k] + e + pad;
            }
            if (!(lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]) || fabsf(lo[0]) > 1e30f || fabsf(hi[0]) > 1e30f ||
                fabsf(lo[1]) > 1e30f || fabsf(hi[1]) > 1e30f || fabsf(lo[2]) > 1e30f || fabsf(hi[2]) > 1e30f) {
#pragma unroll
                for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = FLT_MAX; }
            }
        }
    } else if (i < P_pad) {
        SurfelRec r;
        r.r0 = make_float4(0.f, 0.f, 0.f, -1.0f); r.r1 = make_float4(0.f, 0.f, 0.f, 0.f);
        r.r2 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); r.r3 = make_float4(0.f, 0.f, 1.f, 0.f);
        rec[i] = r;
    }
    const int c = i & 7;
    if (i < P_pad) {
        Node8& n = leaf[i >> 3];
        n.lox[c] = lo[0]; n.loy[c] = lo[1]; n.loz[c] = lo[2];
        n.hix[c] = hi[0]; n.hiy[c] = hi[1]; n.hiz[c] = hi[2];
    }
    // compact copy: bounds of the 8 surfels of this leaf (the 8 threads are consecutive lanes), then 8-bit boxes
    const bool empty = lo[0] == FLT_MAX;
    LeafQ& lq = leafq[min(i, P_pad - 1) >> 3];                       // written only by threads i < P_pad
    float ul[3], uh[3];                                              // union of the leaf's 8 boxes (l > h: all empty)
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float l = empty ? FLT_MAX : lo[k], h = empty ? -FLT_MAX : hi[k];
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        ul[k] = l; uh[k] = h;
        const bool none = l > h;                                     // all 8 slots empty
        const float base = none ? FLT_MAX : l;
        const float sc = none ? 0.0f : fmaxf((h - l) * (1.0f / 255.0f), 1e-30f);
        int ql = 255, qh = 0;                                        // empty slot: inverted byte box (decodes inside the leaf box,
        if (!empty) {                                                // its surfel fails the exact test)
            ql = (int)floorf((lo[k] - base) / sc); ql = min(max(ql, 0), 255);
            while (ql > 0 && fmaf((float)ql, sc, base) > lo[k]) ql--;
            qh = (int)ceilf((hi[k] - base) / sc); qh = min(max(qh, 0), 255);
            while (qh < 255 && fmaf((float)qh, sc, base) < hi[k]) qh++;
            if (fmaf((float)qh, sc, base) < hi[k]) { ql = 0; qh = 255; }   // cannot happen with sc >= (h-l)/255; belt and braces
        }
        if (i < P_pad) {
            lq.qlo[k][c] = (unsigned char)ql; lq.qhi[k][c] = (unsigned char)qh;
            if (c == 0) { lq.lo[k] = base; lq.sc[k] = sc; }
        }
    }
    // Levels 1 and 2 of the hierarchy, while the boxes are in registers (what k_fit would compute from the leaves it re-reads:
    // unions skip empty slots, an all-empty child is lo = hi = +FLT_MAX). Level 1: child (i >> 3) & 7 of node i >> 6 is the
    // union of leaf i >> 3, which its 8 lanes just formed. Level 2: child (i >> 6) & 7 of node i >> 9 is the union of 64
    // consecutive threads = two warps, met through shared memory.
    if (!lvl1) return;                                               // a single leaf: no upper level (block-uniform)
    if (c == 0 && (i >> 6) < n1) {
        const bool none = ul[0] > uh[0];
        Node8& n = lvl1[i >> 6];
        const int cc = (i >> 3) & 7;
        n.lox[cc] = none ? FLT_MAX : ul[0]; n.loy[cc] = none ? FLT_MAX : ul[1]; n.loz[cc] = none ? FLT_MAX : ul[2];
        n.hix[cc] = none ? FLT_MAX : uh[0]; n.hiy[cc] = none ? FLT_MAX : uh[1]; n.hiz[cc] = none ? FLT_MAX : uh[2];
    }
    if (!lvl2) return;
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {
            ul[k] = fminf(ul[k], __shfl_xor_sync(0xffffffffu, ul[k], o));
            uh[k] = fmaxf(uh[k], __shfl_xor_sync(0xffffffffu, uh[k], o));
        }
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { s_box[wib][k] = ul[k]; s_box[wib][3 + k] = uh[k]; }
    }
    __syncthreads();
    if (lane == 0 && (wib & 1) == 0 && (i >> 9) < n2) {
        float l[3], h[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { l[k] = fminf(s_box[wib][k], s_box[wib + 1][k]); h[k] = fmaxf(s_box[wib][3 + k], s_box[wib + 1][3 + k]); }
        const bool none = l[0] > h[0];
        Node8& n = lvl2[i >> 9];
        const int cc = (i >> 6) & 7;
        n.lox[cc] = none ? FLT_MAX : l[0]; n.loy[cc] = none ? FLT_MAX : l[1]; n.loz[cc] = none ? FLT_MAX : l[2];
        n.hix[cc] = none ? FLT_MAX : h[0]; n.hiy[cc] = none ? FLT_MAX : h[1]; n.hiz[cc] = none ? FLT_MAX : h[2];
    }
}

// child c of node j of a level >= 1 is the union of the 8 boxes of node 8j+c one level below. The child level may have been
// written earlier in the SAME launch (k_fit_top): its loads bypass L1.
__device__ __forceinline__ void fit_child(int t, int n_child, const Node8* child, Node8* parent)
{
    const int j = t >> 3, c = t & 7, cj = 8 * j + c;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool any = false;
    if (cj < n_child) {
        const float4* ch = reinterpret_cast<const float4*>(child + cj);          // lox[8] loy[8] loz[8] hix[8] hiy[8] hiz[8]
        float v[48];
#pragma unroll
        for (int q = 0; q < 12; q++) { const float4 x = __ldcg(ch + q); v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w; }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (v[k] != FLT_MAX) {                              // skip empty slots
                any = true;
                lo[0] = fminf(lo[0], v[k]); lo[1] = fminf(lo[1], v[8 + k]); lo[2] = fminf(lo[2], v[16 + k]);
                hi[0] = fmaxf(hi[0], v[24 + k]); hi[1] = fmaxf(hi[1], v[32 + k]); hi[2] = fmaxf(hi[2], v[40 + k]);
            }
        }
    }
    if (!any) { lo[0] = lo[1] = lo[2] = FLT_MAX; hi[0] = hi[1] = hi[2] = FLT_MAX; }
    Node8& n = parent[j];
    n.lox[c] = lo[0]; n.loy[c] = lo[1]; n.loz[c] = lo[2];
    n.hix[c] = hi[0]; n.hiy[c] = hi[1]; n.hiz[c] = hi[2];
}

// Every level from `first` up in ONE launch (the six k_fit launches of a 2 M surfel build were 0.077 ms, nearly all of it launch
// latency on levels of a few thousand nodes and fewer). All blocks fit level `first`; the block that finishes last (ticket
// counter, reset for the next build) walks the remaining levels — a few hundred nodes in total — on its own.
struct FitLevels { int first, levels; int off[LRT_MAX_LEVELS]; int cnt[LRT_MAX_LEVELS]; };

__global__ void __launch_bounds__(256) k_fit_top(Node8* nodes, FitLevels fl, unsigned* ticket)
{
    __shared__ bool s_last;
    {
        const int l = fl.first, t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t < fl.cnt[l] * 8) fit_child(t, fl.cnt[l - 1], nodes + fl.off[l - 1], nodes + fl.off[l]);
    }
    if (fl.first + 1 >= fl.levels) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(ticket, 1u);
        s_last = prev == gridDim.x - 1;
        if (s_last) *ticket = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int l = fl.first + 1; l < fl.levels; l++) {
        for (int t = threadIdx.x; t < fl.cnt[l] * 8; t += blockDim.x) fit_child(t, fl.cnt[l - 1], nodes + fl.off[l - 1], nodes + fl.off[l]);
        __syncthreads();
    }
}

} // namespace

int lrt_build_impl(lrt_ctx* ctx, int P, const float* means, const float* scales, const float* rots,
                   const float* opac, float mod, bool refit, cudaStream_t s)
{
    if (P <= 0 || !means || !scales || !rots || !opac) { ctx->set_error("lrt_build: P must be > 0 and arrays non-null"); return LRT_ERR_INVALID; }
    if (!(mod > 0.0f)) { ctx->set_error("lrt_build: scale_modifier must be > 0"); return LRT_ERR_INVALID; }
    if (refit && (!ctx->built || ctx->P != P)) { ctx->set_error("lrt_refit: no structure built for this P (call lrt_build)"); return LRT_ERR_STATE; }
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));

    // level layout
    const int P_pad = (P + 7) & ~7;
    int cnt[LRT_MAX_LEVELS], off[LRT_MAX_LEVELS], L = 0;
    long long total = 0;
    int n = P_pad / 8;
    for (;;) {
        if (L >= LRT_MAX_LEVELS) { ctx->set_error("lrt_build: too many levels"); return LRT_ERR_INVALID; }
        cnt[L] = n; off[L] = (int)total; total += n; L++;
        if (n == 1) break;
        n = (n + 7) / 8;
    }
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->rec, sizeof(SurfelRec) * (size_t)P_pad));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->nodes, sizeof(Node8) * (size_t)total));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->perm_a, sizeof(unsigned) * (size_t)P));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->rec_g, sizeof(SurfelRec) * (size_t)P));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->leafq, sizeof(LeafQ) * (size_t)(P_pad / 8)));
    const int TB = 256;
    if (!refit) {
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->perm_b, sizeof(unsigned) * (size_t)P));
        const bool wide = ctx->opt_morton_bits > 32;
        const size_t ksz = wide ? sizeof(unsigned long long) : sizeof(unsigned);
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->keys_a, ksz * (size_t)P));
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->keys_b, ksz * (size_t)P));
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bounds, sizeof(int) * 64));
        size_t tmp_bytes = 0;
        cub::DoubleBuffer<unsigned> dv((unsigned*)ctx->perm_b.p, (unsigned*)ctx->perm_a.p);
        cub::DoubleBuffer<unsigned> dk32((unsigned*)ctx->keys_a.p, (unsigned*)ctx->keys_b.p);
        cub::DoubleBuffer<unsigned long long> dk64((unsigned long long*)ctx->keys_a.p, (unsigned long long*)ctx->keys_b.p);
        const int bits32 = ctx->opt_morton_bits == 32 ? 32 : 30;
        // 32-bit keys are sorted on their top 24 bits only: 16.7 M cells for a few million surfels, the order inside a cell is
        // irrelevant (the sort is stable), and the radix sort takes three 8-bit passes instead of four
        const int lo_bit = bits32 == 32 ? 32 - ctx->opt_sort_key_bits : 0;
        if (wide) { LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk64, dv, P, 0, 63, s)); }
        else { LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk32, dv, P, lo_bit, bits32, s)); }
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->sort_tmp, tmp_bytes));
        k_bounds_init<<<1, 32, 0, s>>>((int*)ctx->bounds.p);
        const int gb = min((P / 4 + TB - 1) / TB + 1, 148 * 4);
        ctx->span_begin("k_bounds", s); k_bounds<<<gb, TB, 0, s>>>(P, means, (int*)ctx->bounds.p); ctx->span_end(s);
        if (wide) {
            k_morton64<<<(P + TB - 1) / TB, TB, 0, s>>>(P, means, (const int*)ctx->bounds.p, (unsigned long long*)ctx->keys_a.p,
                                                         (unsigned*)ctx->perm_b.p);
            LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp_bytes, dk64, dv, P, 0, 63, s));
        } else {
            if (bits32 == 32) {
                k_morton_plan<<<1, 32, 0, s>>>((const int*)ctx->bounds.p, (int*)ctx->bounds.p + 8);
                k_morton32<<<(P + TB - 1) / TB, TB, 0, s>>>(P, means, (const int*)ctx->bounds.p, (const int*)ctx->bounds.p + 8,
                                                             (unsigned*)ctx->keys_a.p, (unsigned*)ctx->perm_b.p, ctx->opt_sort_key_bits);
            }
            else
                k_morton<<<(P + TB - 1) / TB, TB, 0, s>>>(P, means, (const int*)ctx->bounds.p, (unsigned*)ctx->keys_a.p,
                                                           (unsigned*)ctx->perm_b.p);
            ctx->span_begin("radix_sort", s);
            LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp_bytes, dk32, dv, P, lo_bit, bits32, s));
            ctx->span_end(s);
        }
        ctx->launches += 3 + 2 + (wide ? 8 : 4);     // bounds_init, bounds, morton + radix sort (histogram, scan, one onesweep launch per 8-bit digit)
        if (dv.Current() != (unsigned*)ctx->perm_a.p) {          // keep the permutation in perm_a
            LRT_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->perm_a.p, dv.Current(), sizeof(unsigned) * (size_t)P,
                                              cudaMemcpyDeviceToDevice, s));
        }
    }
    Node8* nodes = (Node8*)ctx->nodes.p;
    // k_records writes levels 0, 1 and 2; its grid covers every child slot of the highest of them
    const int n_thr = L >= 3 ? cnt[2] * 512 : (L == 2 ? cnt[1] * 64 : P_pad);
    ctx->span_begin("k_records", s);
    k_records<<<(n_thr + TB - 1) / TB, TB, 0, s>>>(P, P_pad, (const unsigned*)ctx->perm_a.p, means, scales, rots, opac, mod, (SurfelRec*)ctx->rec.p,
                                                  nodes + off[0], (SurfelRec*)ctx->rec_g.p, (LeafQ*)ctx->leafq.p,
                                                  L >= 2 ? nodes + off[1] : nullptr, L >= 2 ? cnt[1] : 0, L >= 3 ? nodes + off[2] : nullptr, L >= 3 ? cnt[2] : 0);
    ctx->span_end(s);
    ctx->launches += 1;
    if (L > 3) {
        if (!ctx->fit_ticket.p) {
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->fit_ticket, 64));
            LRT_CUDA_TRY(ctx, cudaMemsetAsync(ctx->fit_ticket.p, 0, 64, s));
        }
        FitLevels fl;
        fl.first = 3; fl.levels = L;
        for (int l = 0; l < LRT_MAX_LEVELS; l++) { fl.off[l] = l < L ? off[l] : 0; fl.cnt[l] = l < L ? cnt[l] : 0; }
        ctx->span_begin("k_fit", s);
        k_fit_top<<<(cnt[3] * 8 + TB - 1) / TB, TB, 0, s>>>(nodes, fl, (unsigned*)ctx->fit_ticket.p);
        ctx->span_end(s);
        ctx->launches += 1;
    }
    LRT_CUDA_TRY(ctx, cudaGetLastError());

    ctx->P = P; ctx->P_pad = P_pad; ctx->levels = L; ctx->n_nodes = total; ctx->scale_modifier = mod;
    for (int i = 0; i < LRT_MAX_LEVELS; i++) { ctx->level_off[i] = i < L ? off[i] : 0; ctx->level_cnt[i] = i < L ? cnt[i] : 0; }
    ctx->built = true;
    if (refit) ctx->refits++; else ctx->builds++;
    return LRT_OK;
}
