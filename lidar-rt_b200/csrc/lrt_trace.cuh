// lrt_trace.cuh — one k-buffer round of a ray through the implicit 8-wide LBVH.
//
// Replaces optixTrace + __anyhit__ot of the reference (forward.cu:48-63, :312-356): find the
// (at most) 16 nearest proxy hits of the re-based ray (o', d) with t' in (0, 1e16), ascending.
// Differences in HOW (results are identical):
//   * the reference's any-hit program must see every triangle on the whole ray in every round
//     (optixIgnoreIntersection never shortens the ray); here subtrees whose entry distance exceeds
//     the current 16th-nearest hit are culled, and children are entered nearest-first so that the
//     bound tightens early;
//   * the proxy is one analytic quad |u|,|v| <= f per Gaussian instead of two triangles;
//   * traversal is stackless: the hierarchy is implicit (children of node j at level l are nodes
//     8j..8j+7 at level l-1), so the only state is (level, node) plus one pending-children byte per
//     level packed into a 64-bit trail. A node with >= 2 pending children is re-evaluated when the
//     traversal climbs back to it (its 192 B are L1-resident), which both restores front-to-back
//     order and drops children that fell behind the shrunken bound.
#pragma once

#include "lrt_common.cuh"

#ifdef LRT_STATS
// development statistics build: [0..7] node evaluations per level, [8] quad tests, [9] quad hits,
// [10] re-evaluations (climb-backs), [11] rounds
extern __device__ unsigned long long g_lrt_stats[16];
#define LRT_STAT(i) atomicAdd(&g_lrt_stats[i], 1ull)
#else
#define LRT_STAT(i)
#endif

#define LRT_KEY_EMPTY 0x5A0E1BCAFFFFFFFFull     // (bits(1e16f) << 32) | 0xffffffff : nothing at t' >= 1e16

struct RaySetup {
    float ox, oy, oz;     // re-based origin o' = o + base d (forward.cu:291)
    float dx, dy, dz;
    float ix, iy, iz;     // 1 / d (clamped away from 0) for the slab tests
    float px, py, pz;     // o' * (1/d)
};

struct Trav {
    int level;
    unsigned node;
    unsigned pend;                 // children of (level, node) still to consider
    unsigned long long trail;      // pending-children byte of every ancestor
};

__device__ __forceinline__ float safe_inv(float d)
{
    const float a = fabsf(d) < 1e-18f ? copysignf(1e-18f, d) : d;
    return 1.0f / a;
}

__device__ __forceinline__ void ray_setup(RaySetup& r, const float* o, const float* d, float base)
{
    r.ox = o[0] + base * d[0]; r.oy = o[1] + base * d[1]; r.oz = o[2] + base * d[2];
    r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
    r.ix = safe_inv(d[0]); r.iy = safe_inv(d[1]); r.iz = safe_inv(d[2]);
    r.px = r.ox * r.ix; r.py = r.oy * r.iy; r.pz = r.oz * r.iz;
}

// Analytic proxy test, same operation order as the oracle's quad_hit().
__device__ __forceinline__ bool quad_hit(const SurfelRec* __restrict__ rec, int prim, const RaySetup& r, float& t_out, int& g_out)
{
    const float4 a0 = ld_f4(&rec[prim].r0), a3 = ld_f4(&rec[prim].r3);
    const float c0 = a0.x - r.ox, c1 = a0.y - r.oy, c2 = a0.z - r.oz;
    const float den = a3.x * r.dx + a3.y * r.dy + a3.z * r.dz;
    const float num = a3.x * c0 + a3.y * c1 + a3.z * c2;
    const float t = num / den;
    if (!(t > 0.0f)) return false;
    const float4 a1 = ld_f4(&rec[prim].r1), a2 = ld_f4(&rec[prim].r2);
    const float r0 = (r.ox + t * r.dx) - a0.x, r1 = (r.oy + t * r.dy) - a0.y, r2 = (r.oz + t * r.dz) - a0.z;
    const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;
    const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
    if (!(fabsf(u) <= a0.w && fabsf(v) <= a0.w)) return false;
    t_out = t;
    g_out = __float_as_int(a2.w);
    return t < LRT_TMAX;
}

// Candidate test for the wavefront's hit bins: the same plane hit, but with the quad bounds relaxed so that the bin is a
// SUPERSET of what the exact test accepts from any re-based origin on this ray. The exact quad_hit() decides, per round, which
// candidates are slots.
//
// How far apart can the two computations of a hit's depth be — t from the ray's origin o (the bin key) and t' + base from the
// re-based origin o' = o + base d (what a round computes)? Both evaluate n.(mu - origin) / n.d in fp32; with u = 2^-24,
//   S_c = sum |n_k| |mu_k - o_k|,  S_d = sum |n_k| |d_k|,  S_o = sum |n_k| |o_k|
// a first-order bound of every rounding on both paths (differences, the three-term dot products, the division, o' itself and the
// final t' + base) is   |t - (t' + base)| <= u (8 S_c + 12 |t| S_d + S_o) / |n.d| + 3 u |t|.
// It grows like 1 / |n.d|: a ground surfel hit at a grazing angle of 2 degrees at 50 m, or a scene placed kilometres from the world
// origin, exceeds the fixed millimetre margin the sorted-bin scan used to rely on. The bound (with a factor 2) is
// returned in e_out: it widens this test's limits by the in-plane motion it allows, the beam-grid window (lrt_beamgrid.cuh), and —
// as the per-ray maximum — the margin with which the compositing passes skip and stop in the sorted bin (wf_margin()).
#define LRT_ERR_FLOOR 1e-3f        // at or below the fixed margin the bound changes nothing: no per-ray bookkeeping
#define LRT_ERR_CAP 5e-2f          // the bound is clamped here (window padding, test limits); a candidate at the cap is tested in every round of its ray
__device__ __forceinline__ bool quad_candidate(const SurfelRec* __restrict__ rec, int prim, const RaySetup& r, float& t_out, int& g_out, float& e_out)
{
    const float4 a0 = ld_f4(&rec[prim].r0), a3 = ld_f4(&rec[prim].r3);
    const float c0 = a0.x - r.ox, c1 = a0.y - r.oy, c2 = a0.z - r.oz;
    const float den = a3.x * r.dx + a3.y * r.dy + a3.z * r.dz;
    const float num = a3.x * c0 + a3.y * c1 + a3.z * c2;
    const float t = num / den;
    if (!(t > -1e-3f)) return false;
    const float4 a1 = ld_f4(&rec[prim].r1), a2 = ld_f4(&rec[prim].r2);
    const float anx = fabsf(a3.x), any_ = fabsf(a3.y), anz = fabsf(a3.z);
    const float Sc = anx * fabsf(c0) + any_ * fabsf(c1) + anz * fabsf(c2);
    const float Sd = anx * fabsf(r.dx) + any_ * fabsf(r.dy) + anz * fabsf(r.dz);
    const float So = anx * fabsf(r.ox) + any_ * fabsf(r.oy) + anz * fabsf(r.oz);
    const float at = fabsf(t);
    float e = 1.2e-7f * ((8.0f * Sc + 12.0f * at * Sd + So) / fabsf(den)) + 3.6e-7f * at;          // 2 u (...) / |n.d| + 6 u |t|
    if (!(e < LRT_ERR_CAP)) e = LRT_ERR_CAP;                                                      // also NaN / inf (n.d == 0)
    const float r0 = (r.ox + t * r.dx) - a0.x, r1 = (r.oy + t * r.dy) - a0.y, r2 = (r.oz + t * r.dz) - a0.z;
    const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;
    const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
    const float lim = a0.w + 1e-3f * (1.0f + a0.w);
    // a depth error e moves the hit point by e d: |du| <= e sum |Lu_k d_k|
    const float gu = fabsf(a1.x * r.dx) + fabsf(a1.y * r.dy) + fabsf(a1.z * r.dz), gv = fabsf(a2.x * r.dx) + fabsf(a2.y * r.dy) + fabsf(a2.z * r.dz);
    if (!(fabsf(u) <= lim + e * gu && fabsf(v) <= lim + e * gv)) return false;
    t_out = fmaxf(t, 1e-30f);                       // > 0: key t = 0 is reserved for wild candidates (wf_append)
    g_out = __float_as_int(a2.w);
    e_out = e;
    return t < LRT_TMAX;
}

// margin of the sorted-bin scan around depth t for a ray whose worst candidate error bound is em. The last term covers the
// ORDER of the split passes' candidate stream: bins of up to 64 candidates are sorted on 32-bit keys whose low 5-6 bits carry the
// candidate's slot instead of the depth's last mantissa bits (lrt_split.cuh), so the stream is sorted only up to 7.6e-6 t.
__device__ __forceinline__ float wf_margin(float t, float em) { return fmaxf(1e-3f + 1e-5f * fabsf(t), em) + 1e-5f * fabsf(t); }

// Sorted insertion into the register-resident k-buffer (ascending 64-bit keys = (t' bits, Gaussian id)).
__device__ __forceinline__ void kbuf_insert(unsigned long long (&kb)[LRT_KBUF], unsigned long long key)
{
    if (key >= kb[LRT_KBUF - 1]) return;
#pragma unroll
    for (int i = 0; i < LRT_KBUF; i++) {
        const unsigned long long cur = kb[i];
        const bool sw = key < cur;
        kb[i] = sw ? key : cur;
        key = sw ? cur : key;
    }
}

// 8 slab tests of one node against [0, tmax]: mask of the children in `pend` the ray enters, and the
// one it enters first.
__device__ __forceinline__ unsigned node_eval(const Node8* __restrict__ node, const RaySetup& r, float tmax,
                                              unsigned pend, int& nearest)
{
    const float4* p = reinterpret_cast<const float4*>(node);
    unsigned m = 0;
    float best = 3.0e38f;
    nearest = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const float4 lx = ld_f4(p + h), ly = ld_f4(p + 2 + h), lz = ld_f4(p + 4 + h);
        const float4 hx = ld_f4(p + 6 + h), hy = ld_f4(p + 8 + h), hz = ld_f4(p + 10 + h);
        const float lxs[4] = {lx.x, lx.y, lx.z, lx.w}, lys[4] = {ly.x, ly.y, ly.z, ly.w}, lzs[4] = {lz.x, lz.y, lz.z, lz.w};
        const float hxs[4] = {hx.x, hx.y, hx.z, hx.w}, hys[4] = {hy.x, hy.y, hy.z, hy.w}, hzs[4] = {hz.x, hz.y, hz.z, hz.w};
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float x0 = fmaf(lxs[c], r.ix, -r.px), x1 = fmaf(hxs[c], r.ix, -r.px);
            const float y0 = fmaf(lys[c], r.iy, -r.py), y1 = fmaf(hys[c], r.iy, -r.py);
            const float z0 = fmaf(lzs[c], r.iz, -r.pz), z1 = fmaf(hzs[c], r.iz, -r.pz);
            const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
            const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
            const bool hit = (tn <= tf) && ((pend >> (4 * h + c)) & 1u);
            if (hit) {
                m |= 1u << (4 * h + c);
                if (tn < best) { best = tn; nearest = 4 * h + c; }
            }
        }
    }
    return m;
}

__device__ __forceinline__ void trav_init(const BvhView& bvh, Trav& tv, unsigned long long (&kb)[LRT_KBUF])
{
#pragma unroll
    for (int i = 0; i < LRT_KBUF; i++) kb[i] = LRT_KEY_EMPTY;
    tv.level = bvh.levels - 1; tv.node = 0; tv.pend = 0xffu; tv.trail = 0;
    LRT_STAT(11);
}

// Climb to the nearest ancestor with pending children (or finish). Returns true when the round is complete.
// O(1): the trail bytes of all levels between the current one and that ancestor are zero (a level is
// only left upwards once its byte is consumed), so the target level is the first non-zero byte above.
__device__ __forceinline__ bool trav_climb(const BvhView& bvh, Trav& tv)
{
    const int sh = 8 * (tv.level + 1);
    const unsigned long long up = sh < 64 ? (tv.trail >> sh) : 0ull;
    if (up == 0) { tv.level = bvh.levels; return true; }
    const int k = (__ffsll((long long)up) - 1) >> 3;          // levels to skip above level + 1
    const int lv = tv.level + 1 + k;
    tv.node >>= 3 * (k + 1);
    tv.level = lv;
    const unsigned p = (unsigned)(tv.trail >> (8 * lv)) & 0xffu;
    if ((p & (p - 1)) == 0) {                               // a single child left: enter it directly
        tv.trail &= ~(0xffull << (8 * lv));
        tv.node = tv.node * 8u + (__ffs(p) - 1); tv.level--; tv.pend = 0xffu;
    } else {
        tv.pend = p;                                        // re-evaluated by the next step: order + culling refresh
    }
    return false;
}

// One node evaluation. Returns true when the round's traversal is complete.
__device__ __forceinline__ bool trav_step(const BvhView& bvh, const RaySetup& r, unsigned long long (&kb)[LRT_KBUF], Trav& tv)
{
#ifdef LRT_NO_CULL   // experiment: enumerate every hit on the ray (what the reference's any-hit program sees)
    const float tmax = LRT_TMAX;
#else
    const float tmax = __uint_as_float((unsigned)(kb[LRT_KBUF - 1] >> 32));
#endif
    int nearest;
    LRT_STAT(tv.level); if (tv.pend != 0xffu) { LRT_STAT(10); }
    unsigned m = node_eval(bvh.nodes + bvh.level_off[tv.level] + tv.node, r, tmax, tv.pend, nearest);
    if (tv.level == 0) {
        while (m) {
            const int c = __ffs(m) - 1; m &= m - 1;
            float t; int g;
            LRT_STAT(8);
            if (quad_hit(bvh.rec, (int)(tv.node * 8u + c), r, t, g)) {   // ties in t' resolve by the caller's Gaussian index
                LRT_STAT(9);
                kbuf_insert(kb, ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)g);
            }
        }
    }
    if (m) {                                            // enter the nearest child, remember the others
        m &= ~(1u << nearest);
        tv.trail = (tv.trail & ~(0xffull << (8 * tv.level))) | ((unsigned long long)m << (8 * tv.level));
        tv.node = tv.node * 8u + nearest; tv.level--; tv.pend = 0xffu;
        return false;
    }
    return trav_climb(bvh, tv);
}

// ---- split form of trav_step for the persistent kernel: node phase / leaf phase / navigation ----
// Node phase: evaluate the current node. Returns 0 = keep traversing, 1 = round complete,
// 2 = at a leaf with candidate surfels in `leaf_mask` (to be consumed by trav_leaf_one).
__device__ __forceinline__ int trav_node(const BvhView& bvh, const RaySetup& r, const unsigned long long (&kb)[LRT_KBUF], Trav& tv,
                                         unsigned& leaf_mask)
{
    const float tmax = __uint_as_float((unsigned)(kb[LRT_KBUF - 1] >> 32));
    int nearest;
    LRT_STAT(tv.level); if (tv.pend != 0xffu) { LRT_STAT(10); }
    unsigned m = node_eval(bvh.nodes + bvh.level_off[tv.level] + tv.node, r, tmax, tv.pend, nearest);
    if (tv.level == 0) {
        if (m) { leaf_mask = m; return 2; }
        return trav_climb(bvh, tv) ? 1 : 0;
    }
    if (m) {
        m &= ~(1u << nearest);
        tv.trail = (tv.trail & ~(0xffull << (8 * tv.level))) | ((unsigned long long)m << (8 * tv.level));
        tv.node = tv.node * 8u + nearest; tv.level--; tv.pend = 0xffu;
        return 0;
    }
    return trav_climb(bvh, tv) ? 1 : 0;
}

// Leaf phase: test ONE candidate surfel of the current leaf. Returns like trav_node (2 = more candidates).
__device__ __forceinline__ int trav_leaf_one(const BvhView& bvh, const RaySetup& r, unsigned long long (&kb)[LRT_KBUF], Trav& tv,
                                             unsigned& leaf_mask)
{
    const int c = __ffs(leaf_mask) - 1; leaf_mask &= leaf_mask - 1;
    float t; int g;
    LRT_STAT(8);
    if (quad_hit(bvh.rec, (int)(tv.node * 8u + c), r, t, g)) {
        LRT_STAT(9);
        kbuf_insert(kb, ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)g);
    }
    if (leaf_mask) return 2;
    return trav_climb(bvh, tv) ? 1 : 0;
}

__device__ __forceinline__ int kbuf_count(const unsigned long long (&kb)[LRT_KBUF])
{
    int n = 0;
#pragma unroll
    for (int i = 0; i < LRT_KBUF; i++) n += kb[i] != LRT_KEY_EMPTY;
    return n;
}

// One full round (monolithic kernels). On return `kb` holds the nearest hits ascending; returns how
// many are valid. 16 valid entries <=> the reference's `payload.cnt >= CHUNK_SIZE` (forward.cu:282).
__device__ __forceinline__ int trace_round(const BvhView& bvh, const RaySetup& r, unsigned long long (&kb)[LRT_KBUF]
#ifdef LRT_STATS
                                           , int& node_visits
#endif
                                           )
{
    Trav tv;
    trav_init(bvh, tv, kb);
    for (;;) {
#ifdef LRT_STATS
        node_visits++;
#endif
        if (trav_step(bvh, r, kb, tv)) break;
    }
    return kbuf_count(kb);
}
