// lrt_forward.cu — forward device program: k-buffer rounds + front-to-back compositing.
//
// Replaces __raygen__ot of the reference's forward pipeline
// (submodules/diff-lidar-tracer/optix_tracer/forward.cu:146-308) and TraceSurfelsCUDA
// (trace_surfels.cpp:151-265). One thread per ray; a round = trace_round() (lrt_trace.cuh);
// the compositing loop follows forward.cu:201-292 rule for rule:
//   depth = t' + base; skip depth < 0.2; skip cos == 0; alpha = min(.99, o exp(-(u^2+v^2)/2));
//   skip alpha < 1/255; stop BEFORE the hit that would take T below 1e-4; next round starts at
//   last depth + 1e-5 when the buffer was full.
// Additionally records, per ray, the ordered list of contributing surfels (id, depth) so the
// backward pass can replay it instead of traversing again.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "lrt_ctx.cuh"
#include "lrt_trace.cuh"

namespace {

__device__ __forceinline__ void load_sh(const float* __restrict__ shs, int g, int M, int nb, float* sh)
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

struct FwdArgs;
__device__ __forceinline__ void load_sh_any(const FwdArgs& a, int g, int nb, float* sh);

struct FwdArgs {
    int R; const float* ray_o; int ray_o_stride; const float* ray_d; const float* bg; const float* shs; int D, M;
    const ShTab* sh_tab;         // non-null: SH rows are read in place through the parts table (shs == nullptr)
    float* out; float* accum_w; int32_t* hit_gidx; float* hit_t; float4* hit_aux; int32_t* hit_cnt; int cap; int32_t* slot_cnt;
    int grid_w;                  // > 0: rays form a row-major (R / grid_w, grid_w) range image -> 4 x 8 warp tiles
    int* work_counter;           // persistent kernel: next work slot
};

__device__ __forceinline__ void load_sh_any(const FwdArgs& a, int g, int nb, float* sh)
{
    if (a.sh_tab) load_sh_parts(a.sh_tab, g, nb, sh);
    else load_sh(a.shs, g, a.M, nb, sh);
}
__host__ __device__ __forceinline__ bool sh_rows_aligned(const FwdArgs& a) { return !a.sh_tab && (a.M & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.shs) & 15) == 0); }

// Per-ray compositing state (forward.cu:174-193).
struct FwdRay {
    float o[3], d[3], dirn[3];
    float C0, C1, C2, Dp, W, T, testT, base, dpt;
    int ncontrib, nslots, last, r;
};

__device__ __forceinline__ void fwd_ray_init(FwdRay& q, int r, const FwdArgs& a)
{
    q.r = r;
#pragma unroll
    for (int k = 0; k < 3; k++) { q.o[k] = a.ray_o[(size_t)r * a.ray_o_stride + k]; q.d[k] = a.ray_d[3 * (size_t)r + k]; }
    const float dl = sqrtf(q.d[0] * q.d[0] + q.d[1] * q.d[1] + q.d[2] * q.d[2]);
#pragma unroll
    for (int k = 0; k < 3; k++) q.dirn[k] = q.d[k] / dl;
    q.C0 = q.C1 = q.C2 = q.Dp = q.W = 0.f; q.T = 1.f; q.testT = 1.f; q.base = 0.f; q.dpt = 0.f;
    q.ncontrib = 0; q.nslots = 0; q.last = -1;
}

// Composite one round's sorted hits (forward.cu:201-292). Returns true if the ray needs another round.
__device__ __forceinline__ bool fwd_shade_round(FwdRay& q, const unsigned long long* hits, int n, const BvhView& bvh, const FwdArgs& a)
{
    const int nb = (a.D + 1) * (a.D + 1);
    bool terminated = false;
    for (int i = 0; i < n; i++) {
        const unsigned long long key = hits[i];
        const int g = (int)(unsigned)(key & 0xffffffffull);
        q.nslots++;
        q.dpt = __uint_as_float((unsigned)(key >> 32)) + q.base;                  // forward.cu:212
        if (q.dpt < LRT_MIN_T) continue;                                          // :214
        const float x0 = q.o[0] + q.dpt * q.d[0], x1 = q.o[1] + q.dpt * q.d[1], x2 = q.o[2] + q.dpt * q.d[2];
        // :220-224 — a re-based round can find the previous round's last surfel again at t' ~ +0
        // (the 1e-5 step is below one ulp of the depth beyond 128 m): it must not composite twice
        if (g == q.last) continue;
        q.last = g;
        const float4 a0 = ld_f4(&bvh.rec_g[g].r0), a1 = ld_f4(&bvh.rec_g[g].r1);
        const float4 a2 = ld_f4(&bvh.rec_g[g].r2), a3 = ld_f4(&bvh.rec_g[g].r3);
        const float r0 = x0 - a0.x, r1 = x1 - a0.y, r2 = x2 - a0.z;
        const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;                        // :139
        const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
        const float cosv = -((a0.x - q.o[0]) * a3.x + (a0.y - q.o[1]) * a3.y + (a0.z - q.o[2]) * a3.z);
        if (cosv == 0.0f) continue;                                               // :233-237
        const float rho = u * u + v * v;
        const float power = -0.5f * rho;
        if (power > 0.0f) continue;
        const float G = expf(power);
        const float alpha = fminf(LRT_ALPHA_MAX, a1.w * G);                       // :249
        if (alpha < 1.0f / 255.0f) continue;
        q.testT = q.T * (1.0f - alpha);
        if (q.testT < LRT_T_MIN) { terminated = true; break; }                    // :253-257
        const float w = alpha * q.T;
        float c[3];
        if (sh_rows_aligned(a)) {
            sh_colour_stream(a.D, q.dirn, a.shs + (size_t)g * a.M * 3, c);        // same sums in the same order, no 48-float staging
        } else {
            float sh[48]; bool cl;
            load_sh_any(a, g, nb, sh);
            sh_colour<false>(a.D, q.dirn, sh, c, cl, nullptr);
        }
        q.C0 += w * c[0]; q.C1 += w * c[1]; q.C2 += w * c[2];
        q.Dp += w * q.dpt; q.W += w;
        if (a.accum_w) atomicAdd(a.accum_w + g, w);                               // :272 (nullptr: redone ray, weights already accumulated)
        if (a.hit_gidx != nullptr && q.ncontrib < a.cap) {
            a.hit_gidx[(size_t)q.ncontrib * a.R + q.r] = g;
            a.hit_t[(size_t)q.ncontrib * a.R + q.r] = q.dpt;
            if (a.hit_aux) a.hit_aux[(size_t)q.ncontrib * a.R + q.r] = make_float4(alpha, c[0], c[1], c[2]);
        }
        q.ncontrib++;
        q.T = q.testT;
    }
    if (terminated || q.testT < LRT_T_MIN || n < LRT_KBUF) return false;          // :282-285
    q.base = (float)((double)q.dpt + LRT_STEP_EPS);                               // :288
    return true;
}

__device__ __forceinline__ void fwd_write(const FwdRay& q, const FwdArgs& a, int node_visits)
{
    float* op = a.out + (size_t)LRT_NCH * q.r;
    op[0] = q.C0 + q.T * a.bg[0]; op[1] = q.C1 + q.T * a.bg[1]; op[2] = q.C2 + q.T * a.bg[2];      // :296-305
    op[3] = q.Dp; op[4] = q.W; op[5] = 0.f; op[6] = 0.f; op[7] = 0.f; op[8] = q.T;
    if (a.hit_cnt) a.hit_cnt[q.r] = q.ncontrib;
#ifdef LRT_STATS
    if (a.slot_cnt) a.slot_cnt[q.r] = (q.nslots & 0xffff) | (min(node_visits, 32767) << 16);   // development statistics build
#else
    (void)node_visits;
    if (a.slot_cnt) a.slot_cnt[q.r] = q.nslots;
#endif
}

// Work slot -> ray index. With a known range-image width, consecutive slots walk 4 x 8 tiles so that
// the 32 rays of a warp are spatial neighbours (shared nodes, similar traversal length); -1 = padding.
__device__ __forceinline__ int slot_to_ray(int s, int R, int grid_w)
{
    if (grid_w <= 0) return s < R ? s : -1;
    const int H = R / grid_w, tiles_x = (grid_w + 7) >> 3;
    const int tile = s >> 5, l = s & 31;
    const int h = (tile / tiles_x) * 4 + (l >> 3), w = (tile % tiles_x) * 8 + (l & 7);
    return (h < H && w < grid_w) ? h * grid_w + w : -1;
}

__host__ __device__ __forceinline__ int num_slots(int R, int grid_w)
{
    if (grid_w <= 0) return R;
    const int H = R / grid_w;
    return ((H + 3) >> 2) * ((grid_w + 7) >> 3) * 32;
}

// ---- kernel A: one thread per ray, whole ray in one go (simple; kept as the comparison baseline and as the
// fallback of the wavefront path)
__device__ __forceinline__ void forward_one_ray(const BvhView& bvh, const FwdArgs& a, int r)
{
    FwdRay q;
    fwd_ray_init(q, r, a);
    int node_visits = 0;
    for (;;) {
        RaySetup rs;
        ray_setup(rs, q.o, q.d, q.base);
        unsigned long long kb[LRT_KBUF];
#ifdef LRT_STATS
        const int n = trace_round(bvh, rs, kb, node_visits);
#else
        const int n = trace_round(bvh, rs, kb);
#endif
        unsigned long long hits[LRT_KBUF];
#pragma unroll
        for (int i = 0; i < LRT_KBUF; i++) hits[i] = kb[i];
        if (!fwd_shade_round(q, hits, n, bvh, a)) break;
#ifdef LRT_NO_CULL
        break;       // experiment build: one un-culled enumeration per ray, statistics only
#endif
    }
    fwd_write(q, a, node_visits);
}

__global__ void __launch_bounds__(128) k_forward(BvhView bvh, FwdArgs a)
{
    const int r = slot_to_ray(blockIdx.x * blockDim.x + threadIdx.x, a.R, a.grid_w);
    if (r < 0) return;
    forward_one_ray(bvh, a, r);
}

// ---- kernel B: persistent threads. Rays differ 20x in traversal length (sky vs. long grazing rays), so
// with one ray per thread a warp idles at ~30 % lane utilisation waiting for its longest ray. Here every
// lane runs a small state machine and pulls a new ray the moment its own finishes:
//   FETCH -> (TRAV: one node evaluation per loop trip)* -> SHADE (16 sorted hits) -> TRAV | FETCH
// All lanes in TRAV execute the same node-evaluation code each trip; shading is batched (it runs when
// >= LRT_SHADE_BATCH lanes wait for it or nobody can traverse) so that its divergent code is amortised.
#define LRT_ST_FETCH 0
#define LRT_ST_TRAV 1
#define LRT_ST_LEAF 2
#define LRT_ST_SHADE 3
#define LRT_ST_DONE 4
#ifndef LRT_SHADE_BATCH
#define LRT_SHADE_BATCH 8
#endif
#ifndef LRT_MIN_LANES
#define LRT_MIN_LANES 5
#endif

#ifndef LRT_FWD_MIN_BLOCKS
#define LRT_FWD_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(128, LRT_FWD_MIN_BLOCKS) k_forward_persistent(BvhView bvh, FwdArgs a)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int S = num_slots(a.R, a.grid_w);
    int st = LRT_ST_FETCH;
    FwdRay q;
    RaySetup rs;
    Trav tv;
    unsigned long long kb[LRT_KBUF];
    unsigned leaf_mask = 0;
    int node_visits = 0;
    q.r = -1;
    for (;;) {
        // 1. refill idle lanes (one atomic per warp)
        const unsigned need = __ballot_sync(FULL, st == LRT_ST_FETCH);
        if (need) {
            int base_slot = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base_slot = atomicAdd(a.work_counter, __popc(need));
            base_slot = __shfl_sync(FULL, base_slot, leader);
            if (st == LRT_ST_FETCH) {
                const int s = base_slot + __popc(need & ((1u << lane) - 1u));
                if (s >= S) st = LRT_ST_DONE;
                else {
                    const int r = slot_to_ray(s, a.R, a.grid_w);
                    if (r >= 0) {
                        fwd_ray_init(q, r, a);
                        ray_setup(rs, q.o, q.d, q.base);
                        trav_init(bvh, tv, kb);
                        node_visits = 0;
                        st = LRT_ST_TRAV;
                    }                                   // padding slot: fetch again next trip
                }
            }
        }
        // 2. what is waiting? Every trip runs the node-evaluation block for lanes in TRAV and the surfel-test
        //    block for lanes in LEAF, so every traversing lane advances one step per trip; a block is skipped
        //    for a trip only when very few lanes want it while many want the other. Shading (long, divergent)
        //    is batched.
        const unsigned wt = __ballot_sync(FULL, st == LRT_ST_TRAV);
        const unsigned wl = __ballot_sync(FULL, st == LRT_ST_LEAF);
        const unsigned ws = __ballot_sync(FULL, st == LRT_ST_SHADE);
        if ((wt | wl | ws) == 0) {
            if (__all_sync(FULL, st == LRT_ST_DONE)) break;
            continue;                                   // only FETCH lanes left (padding slots)
        }
        const int nt = __popc(wt), nl = __popc(wl), ns = __popc(ws);
        if (ns >= LRT_SHADE_BATCH || (nt == 0 && nl == 0)) {
            if (st == LRT_ST_SHADE) {
                unsigned long long hits[LRT_KBUF];
#pragma unroll
                for (int i = 0; i < LRT_KBUF; i++) hits[i] = kb[i];
                const int n = kbuf_count(kb);
                if (fwd_shade_round(q, hits, n, bvh, a)) {
                    ray_setup(rs, q.o, q.d, q.base);
                    trav_init(bvh, tv, kb);
                    st = LRT_ST_TRAV;
                } else {
                    fwd_write(q, a, node_visits);
                    st = LRT_ST_FETCH;
                }
            }
            continue;
        }
        const bool run_leaf = nl > 0 && (nl >= LRT_MIN_LANES || nt < LRT_MIN_LANES);
        const bool run_node = nt > 0 && (nt >= LRT_MIN_LANES || nl < LRT_MIN_LANES);
        const int st_in = st;
        if (run_leaf && st_in == LRT_ST_LEAF) {
            const int res = trav_leaf_one(bvh, rs, kb, tv, leaf_mask);
            st = res == 2 ? LRT_ST_LEAF : (res == 1 ? LRT_ST_SHADE : LRT_ST_TRAV);
        }
        if (run_node && st_in == LRT_ST_TRAV) {
            node_visits++;
            const int res = trav_node(bvh, rs, kb, tv, leaf_mask);
            st = res == 2 ? LRT_ST_LEAF : (res == 1 ? LRT_ST_SHADE : LRT_ST_TRAV);
        }
    }
}

// ---- kernel C ("G8"): 8 lanes per ray, 4 rays per warp, persistent with per-group refill.
// The node is 8 wide, so lane c of a group owns child c: a node evaluation is ONE slab test per lane
// (instead of 8 per lane), the 8 surfels of a leaf are tested in parallel, the 16-entry k-buffer lives
// 2 entries per lane (sorted insertion = two ballots + three shuffles), and a round's 16 hits are
// shaded 8 at a time before an in-order fold composites them. Per-lane state shrinks from ~170 to
// ~90 registers (twice the resident warps) and the dependent chain of a traversal step from ~200
// to ~40 instructions. Arithmetic per hit is unchanged, so results are bit-identical to kernels A/B.
#ifndef LRT_G8_MIN_BLOCKS
#define LRT_G8_MIN_BLOCKS 6
#endif
#define G8_ST_FETCH 0
#define G8_ST_TRAV 1
#define G8_ST_SHADE 2
#define G8_ST_DONE 3
#define G8_F_DPT_OK 1u
#define G8_F_OK 2u

struct G8Slot { float dpt, alpha, c0, c1, c2; int g; unsigned flags; };

// shade ONE k-buffer slot (everything of forward.cu:207-266 that does not depend on earlier slots)
__device__ __forceinline__ void g8_shade_slot(unsigned long long key, bool valid, const FwdRay& q, const BvhView& bvh, const FwdArgs& a,
                                              G8Slot& o)
{
    o.flags = 0; o.dpt = 0.f; o.alpha = 0.f; o.c0 = o.c1 = o.c2 = 0.f; o.g = -1;
    if (!valid) return;
    const int g = (int)(unsigned)(key & 0xffffffffull);
    o.g = g;
    o.dpt = __uint_as_float((unsigned)(key >> 32)) + q.base;                      // forward.cu:212
    if (o.dpt < LRT_MIN_T) return;                                                // :214
    o.flags |= G8_F_DPT_OK;
    const float x0 = q.o[0] + o.dpt * q.d[0], x1 = q.o[1] + o.dpt * q.d[1], x2 = q.o[2] + o.dpt * q.d[2];
    const float4 a0 = ld_f4(&bvh.rec_g[g].r0), a1 = ld_f4(&bvh.rec_g[g].r1);
    const float4 a2 = ld_f4(&bvh.rec_g[g].r2), a3 = ld_f4(&bvh.rec_g[g].r3);
    const float r0 = x0 - a0.x, r1 = x1 - a0.y, r2 = x2 - a0.z;
    const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;                            // :139
    const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
    const float cosv = -((a0.x - q.o[0]) * a3.x + (a0.y - q.o[1]) * a3.y + (a0.z - q.o[2]) * a3.z);
    if (cosv == 0.0f) return;                                                     // :233-237
    const float rho = u * u + v * v;
    const float power = -0.5f * rho;
    if (power > 0.0f) return;
    const float G = expf(power);
    o.alpha = fminf(LRT_ALPHA_MAX, a1.w * G);                                     // :249
    if (o.alpha < 1.0f / 255.0f) return;
    o.flags |= G8_F_OK;
    float c[3];
    if (sh_rows_aligned(a)) {
        sh_colour_stream(a.D, q.dirn, a.shs + (size_t)g * a.M * 3, c);
    } else {
        const int nb = (a.D + 1) * (a.D + 1);
        float sh[48]; bool cl;
        load_sh_any(a, g, nb, sh);
        sh_colour<false>(a.D, q.dirn, sh, c, cl, nullptr);
    }
    o.c0 = c[0]; o.c1 = c[1]; o.c2 = c[2];
}

__global__ void __launch_bounds__(128, LRT_G8_MIN_BLOCKS) k_forward_g8(BvhView bvh, FwdArgs a)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7, gsh = 8 * grp;
    const unsigned gmask = 0xffu << gsh;
    const int S = num_slots(a.R, a.grid_w);
    int st = G8_ST_FETCH;                         // group-uniform
    FwdRay q;                                     // replicated on the 8 lanes of the group
    RaySetup rs;
    Trav tv;
    unsigned long long k0 = LRT_KEY_EMPTY, k1 = LRT_KEY_EMPTY;   // k-buffer entries [sub] and [sub + 8]
    float tmax = LRT_TMAX;                        // t' of entry 15 (replicated)
    int node_visits = 0;
    q.r = -1;
    for (;;) {
        __syncwarp(FULL);
        // 1. refill idle groups (one atomic per warp)
        const unsigned need = __ballot_sync(FULL, st == G8_ST_FETCH);
        if (need) {
            int base_slot = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base_slot = atomicAdd(a.work_counter, __popc(need) >> 3);
            base_slot = __shfl_sync(FULL, base_slot, leader);
            if (st == G8_ST_FETCH) {
                const int s = base_slot + (__popc(need & ((1u << gsh) - 1u)) >> 3);
                if (s >= S) st = G8_ST_DONE;
                else {
                    const int r = slot_to_ray(s, a.R, a.grid_w);
                    if (r >= 0) {
                        fwd_ray_init(q, r, a);
                        ray_setup(rs, q.o, q.d, q.base);
                        tv.level = bvh.levels - 1; tv.node = 0; tv.pend = 0xffu; tv.trail = 0;
                        k0 = k1 = LRT_KEY_EMPTY; tmax = LRT_TMAX;
                        node_visits = 0;
                        LRT_STAT(11);
                        st = G8_ST_TRAV;
                    }
                }
            }
        }
        const unsigned wt = __ballot_sync(FULL, st == G8_ST_TRAV);
        const unsigned ws = __ballot_sync(FULL, st == G8_ST_SHADE);
        if ((wt | ws) == 0) {
            if (__all_sync(FULL, st == G8_ST_DONE)) break;
            continue;
        }
        // 2. traversal step: lane `sub` owns child `sub` of the group's current node
        if (st == G8_ST_TRAV) {
            node_visits++;
            LRT_STAT(tv.level); if (tv.pend != 0xffu) { LRT_STAT(10); }
            const Node8* nd = bvh.nodes + bvh.level_off[tv.level] + tv.node;
            const float lx = ld_f(nd->lox + sub), ly = ld_f(nd->loy + sub), lz = ld_f(nd->loz + sub);
            const float hx = ld_f(nd->hix + sub), hy = ld_f(nd->hiy + sub), hz = ld_f(nd->hiz + sub);
            const float x0 = fmaf(lx, rs.ix, -rs.px), x1 = fmaf(hx, rs.ix, -rs.px);
            const float y0 = fmaf(ly, rs.iy, -rs.py), y1 = fmaf(hy, rs.iy, -rs.py);
            const float z0 = fmaf(lz, rs.iz, -rs.pz), z1 = fmaf(hz, rs.iz, -rs.pz);
            const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
            const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
            const bool hit = (tn <= tf) && ((tv.pend >> sub) & 1u);
            unsigned m = (__ballot_sync(gmask, hit) >> gsh) & 0xffu;
            bool done = false;
            if (tv.level == 0) {
                if (m) {                                                   // the leaf's candidate surfels, in parallel
                    float t = 0.f; int g = 0; bool qh = false;
                    if (hit) { LRT_STAT(8); qh = quad_hit(bvh.rec, (int)(tv.node * 8u + sub), rs, t, g); if (qh) { LRT_STAT(9); } }
                    unsigned qm = (__ballot_sync(gmask, qh) >> gsh) & 0xffu;
                    while (qm) {                                           // distributed sorted insertion, one hit at a time
                        const int src = __ffs(qm) - 1; qm &= qm - 1;
                        const unsigned xh = __shfl_sync(gmask, __float_as_uint(t), src, 8);
                        const unsigned xl = __shfl_sync(gmask, (unsigned)g, src, 8);
                        const unsigned long long x = ((unsigned long long)xh << 32) | xl;
                        const int pos = __popc(__ballot_sync(gmask, k0 < x) & gmask) + __popc(__ballot_sync(gmask, k1 < x) & gmask);
                        const unsigned long long up0 = __shfl_up_sync(gmask, k0, 1, 8);
                        const unsigned long long up1 = __shfl_up_sync(gmask, k1, 1, 8);
                        const unsigned long long k0_7 = __shfl_sync(gmask, k0, 7, 8);
                        if (pos < LRT_KBUF) {
                            k0 = sub < pos ? k0 : (sub == pos ? x : up0);
                            k1 = sub + 8 < pos ? k1 : (sub + 8 == pos ? x : (sub == 0 ? k0_7 : up1));
                        }
                        tmax = __uint_as_float(__shfl_sync(gmask, (unsigned)(k1 >> 32), 7, 8));
                    }
                }
                done = trav_climb(bvh, tv);
            } else if (m) {
                // nearest entered child: redux-min over (tn bits | child) of the entering lanes
                const unsigned key = hit ? ((__float_as_uint(tn) & ~7u) | (unsigned)sub) : 0xffffffffu;
                const int c = (int)(__reduce_min_sync(gmask, key) & 7u);
                m &= ~(1u << c);
                tv.trail = (tv.trail & ~(0xffull << (8 * tv.level))) | ((unsigned long long)m << (8 * tv.level));
                tv.node = tv.node * 8u + c; tv.level--; tv.pend = 0xffu;
            } else {
                done = trav_climb(bvh, tv);
            }
            if (done) st = G8_ST_SHADE;
        }
        // 3. shading: when at least two groups wait for it, or nothing else can run
        const unsigned ws2 = __ballot_sync(FULL, st == G8_ST_SHADE);
        const unsigned wt2 = __ballot_sync(FULL, st == G8_ST_TRAV);
        if (st == G8_ST_SHADE && (__popc(ws2) >= 16 || wt2 == 0)) {
            const int n = __popc(__ballot_sync(gmask, k0 != LRT_KEY_EMPTY) & gmask) + __popc(__ballot_sync(gmask, k1 != LRT_KEY_EMPTY) & gmask);
            G8Slot sa, sb;
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                G8Slot t_;
                g8_shade_slot(pass ? k1 : k0, sub + 8 * pass < n, q, bvh, a, t_);
                if (pass) sb = t_; else sa = t_;
            }
            bool terminated = false;
            for (int i = 0; i < n; i++) {                                   // in-order fold (forward.cu:201-280)
                const int src = i & 7; const bool second = i >= 8;
                const float dpt_i = __shfl_sync(gmask, second ? sb.dpt : sa.dpt, src, 8);
                const unsigned fl_i = __shfl_sync(gmask, second ? sb.flags : sa.flags, src, 8);
                q.nslots++;
                q.dpt = dpt_i;
                if (!(fl_i & G8_F_DPT_OK)) continue;                                          // :214
                const int g_i = __shfl_sync(gmask, second ? sb.g : sa.g, src, 8);
                if (g_i == q.last) continue;                                                  // :220-224
                q.last = g_i;
                if (!(fl_i & G8_F_OK)) continue;
                const float alpha = __shfl_sync(gmask, second ? sb.alpha : sa.alpha, src, 8);
                q.testT = q.T * (1.0f - alpha);
                if (q.testT < LRT_T_MIN) { terminated = true; break; }                        // :253-257
                const float w = alpha * q.T;
                const float c0 = __shfl_sync(gmask, second ? sb.c0 : sa.c0, src, 8);
                const float c1 = __shfl_sync(gmask, second ? sb.c1 : sa.c1, src, 8);
                const float c2 = __shfl_sync(gmask, second ? sb.c2 : sa.c2, src, 8);
                q.C0 += w * c0; q.C1 += w * c1; q.C2 += w * c2;
                q.Dp += w * dpt_i; q.W += w;
                if (sub == src) {
                    atomicAdd(a.accum_w + g_i, w);                                            // :272
                    if (a.hit_gidx != nullptr && q.ncontrib < a.cap) {
                        a.hit_gidx[(size_t)q.ncontrib * a.R + q.r] = g_i;
                        a.hit_t[(size_t)q.ncontrib * a.R + q.r] = dpt_i;
                        if (a.hit_aux) a.hit_aux[(size_t)q.ncontrib * a.R + q.r] = make_float4(alpha, c0, c1, c2);
                    }
                }
                q.ncontrib++;
                q.T = q.testT;
            }
            if (!(terminated || q.testT < LRT_T_MIN || n < LRT_KBUF)) {                       // :282-291
                q.base = (float)((double)q.dpt + LRT_STEP_EPS);
                ray_setup(rs, q.o, q.d, q.base);
                tv.level = bvh.levels - 1; tv.node = 0; tv.pend = 0xffu; tv.trail = 0;
                k0 = k1 = LRT_KEY_EMPTY; tmax = LRT_TMAX;
                LRT_STAT(11);
                st = G8_ST_TRAV;
            } else {
                if (sub == 0) fwd_write(q, a, node_visits);
                st = G8_ST_FETCH;
            }
        }
    }
}

#include "lrt_wavefront.cuh"
#include "lrt_split.cuh"
#include "lrt_beamgrid.cuh"

} // namespace

int lrt_forward_impl(lrt_ctx* ctx, int R, const float* ray_o, int ray_o_stride, const float* ray_d,
                     const float* bg, int P, const float* means, const float* scales, const float* rots,
                     const float* opac, const float* shs, int D, int M, float mod,
                     float* out, float* accum_w, int32_t* hit_gidx, float* hit_t, float* hit_aux, int32_t* hit_cnt,
                     int cap, int32_t* slot_cnt, cudaStream_t s)
{
    if (!ctx->built) { ctx->set_error("lrt_forward: no acceleration structure (call lrt_build first)"); return LRT_ERR_STATE; }
    if (P != ctx->P) { ctx->set_error("lrt_forward: P differs from the built structure"); return LRT_ERR_STATE; }
    if (mod != ctx->scale_modifier) { ctx->set_error("lrt_forward: scale_modifier differs from the built structure"); return LRT_ERR_STATE; }
    if (R == 0 && accum_w) {                                   // empty ray set: only the per-Gaussian weights exist
        LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        LRT_CUDA_TRY(ctx, cudaMemsetAsync(accum_w, 0, sizeof(float) * (size_t)P, s));
        return LRT_OK;
    }
    const ShTab* sh_tab = nullptr;
    if (!shs) {                                                // SH rows in place (lrt_set_sh_parts)
        if (!ctx->sh_parts_n || ctx->sh_parts_P != P || ctx->sh_parts_M != M) { ctx->set_error("lrt_forward: shs is NULL and no matching SH parts are bound (lrt_set_sh_parts)"); return LRT_ERR_STATE; }
        sh_tab = (const ShTab*)ctx->sh_tab.p;
    }
    if (R < 0 || !ray_d || !ray_o || !bg || !out || !accum_w) { ctx->set_error("lrt_forward: null argument"); return LRT_ERR_INVALID; }
    if (ray_o_stride != 0 && ray_o_stride != 3) { ctx->set_error("lrt_forward: ray_o_stride must be 0 or 3"); return LRT_ERR_INVALID; }
    if (D < 0 || D > 3 || M < (D + 1) * (D + 1)) { ctx->set_error("lrt_forward: need 0 <= D <= 3 and M >= (D+1)^2"); return LRT_ERR_INVALID; }
    if ((hit_gidx == nullptr) != (hit_t == nullptr) || (hit_gidx && (cap <= 0 || !hit_cnt))) {
        ctx->set_error("lrt_forward: hit_gidx, hit_t and hit_cnt go together and need cap > 0"); return LRT_ERR_INVALID;
    }
    if (hit_aux && (!hit_gidx || (reinterpret_cast<uintptr_t>(hit_aux) & 15))) { ctx->set_error("lrt_forward: hit_aux needs the hit lists and 16-byte alignment"); return LRT_ERR_INVALID; }
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LRT_CUDA_TRY(ctx, cudaMemsetAsync(accum_w, 0, sizeof(float) * (size_t)P, s));
    if (R == 0) return LRT_OK;
    FwdArgs a;
    a.R = R; a.ray_o = ray_o; a.ray_o_stride = ray_o_stride; a.ray_d = ray_d; a.bg = bg; a.shs = shs; a.D = D; a.M = M; a.sh_tab = sh_tab;
    a.out = out; a.accum_w = accum_w; a.hit_gidx = hit_gidx; a.hit_t = hit_t; a.hit_aux = reinterpret_cast<float4*>(hit_aux); a.hit_cnt = hit_cnt; a.cap = cap; a.slot_cnt = slot_cnt;
    a.grid_w = (ctx->opt_ray_grid_w > 0 && R % ctx->opt_ray_grid_w == 0) ? ctx->opt_ray_grid_w : 0;
    a.work_counter = nullptr;
    const int TB = 128;
    const int S = num_slots(R, a.grid_w);
    if (ctx->opt_forward_kernel == 0) {
        ctx->span_begin("k_forward", s); k_forward<<<(S + TB - 1) / TB, TB, 0, s>>>(ctx->view(), a); ctx->span_end(s);
    } else if (ctx->opt_forward_kernel >= 3) {
        // candidate bins per ray, filled either by the shared-origin beam grid (kernel 4: one pass over the surfel
        // records) or by the breadth-first wavefront through the hierarchy (kernel 3, and any frame with per-ray
        // origins); then per-ray sort + compositing
        const bool beam = ctx->opt_forward_kernel == 4 && ray_o_stride == 0;
        WfBufs w;
        const long long cap_items = (long long)R * 160 + 1024;
        if (cap_items > 0x3fffffffLL) { ctx->set_error("lrt_forward: too many rays for the wavefront work lists"); return LRT_ERR_INVALID; }
        if (!beam) {
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_rs, sizeof(RaySetup) * (size_t)R));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_list_a, sizeof(uint2) * (size_t)cap_items));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_list_b, sizeof(uint2) * (size_t)cap_items));
        }
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_hit_count, sizeof(int) * 5 * (size_t)R));
        // Bin capacity. The split passes take candidates beyond a bin through the overflow list (WfBufs::ov_pairs), so their bins
        // only need to hold the usual ray: 16384 entries for small patches, halved while the array is above 1.5 GB (1024 at one
        // Waymo frame: 1.4 GB; a full street-scene ray carries ~40 candidates, a grazing one several hundred). The other
        // compositing forms drop a ray beyond its bin to the per-ray fallback and keep the 6 GB limit (4096 at one Waymo frame).
        const bool ov_active = ctx->opt_wavefront_shade == 3;
        int hcap = 2 * WF_HCAP_MAX;
        while (hcap > WF_HCAP && (size_t)R * hcap * sizeof(unsigned long long) > (ov_active ? (size_t)3 << 29 : (size_t)6 << 30)) hcap >>= 1;
        if (ctx->opt_bin_cap > 0) hcap = ctx->opt_bin_cap;             // LRT_OPT_BIN_CAP (tests: force the overflow route)
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_bins, sizeof(unsigned long long) * (size_t)R * hcap));
        w.hcap = hcap;
        w.ov_pairs = nullptr; w.ov_pair_cap = 0; w.ov_area = nullptr; w.ov_area_cap = 0; w.ov_base = nullptr; w.ov_fill = nullptr;
        if (ov_active) {
            w.ov_pair_cap = 1 << 22; w.ov_area_cap = 1 << 23;         // 4 M pairs (64 MB), 8 M keys (64 MB)
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_ov_pairs, sizeof(uint4) * (size_t)w.ov_pair_cap));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_ov_area, sizeof(unsigned long long) * (size_t)w.ov_area_cap));
            w.ov_pairs = (uint4*)ctx->wf_ov_pairs.p; w.ov_area = (unsigned long long*)ctx->wf_ov_area.p;
        }
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_fb, sizeof(int) * (size_t)R * 3));
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_ids, sizeof(int) * (size_t)R * 2));
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_keys, sizeof(int) * (size_t)R));
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->counter, sizeof(int) * 16));
        LRT_CUDA_TRY(ctx, cudaMemsetAsync(ctx->counter.p, 0, sizeof(int) * 16, s));
        w.rs = (RaySetup*)ctx->wf_rs.p; w.list_a = (uint2*)ctx->wf_list_a.p; w.list_b = (uint2*)ctx->wf_list_b.p;
        w.cap_items = (int)cap_items; w.counts = (int*)ctx->counter.p; w.hit_count = (int*)ctx->wf_hit_count.p; w.emax = w.hit_count + R; w.nwild = w.hit_count + 2 * (size_t)R;
        if (ov_active) { w.ov_base = w.hit_count + 3 * (size_t)R; w.ov_fill = w.hit_count + 4 * (size_t)R; }
        w.bins = (unsigned long long*)ctx->wf_bins.p; w.fb_list = (int*)ctx->wf_fb.p; w.big_list = (int*)ctx->wf_fb.p + R;
        w.ov_list = (int*)ctx->wf_fb.p + 2 * (size_t)R;
        w.ray_ids = (int*)ctx->wf_ids.p; w.order = nullptr;
        if (ctx->num_sms == 0) {
            int sms = 0;
            LRT_CUDA_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
            ctx->num_sms = sms > 0 ? sms : 148;
        }
        const BvhView bv = ctx->view();
        const int G = ctx->num_sms * 8;                                 // grid-stride kernels: 8 blocks of 256 threads per SM
        if (beam) {
            BgBufs b;
            b.ncell_cap = 2 * R + 2 * BG_MAX_NA;
            const size_t ncell = (size_t)b.ncell_cap + 1;
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bg_ang, sizeof(float2) * (size_t)R));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bg_cell_of, sizeof(int) * (size_t)R));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bg_cells, sizeof(int) * 2 * ncell));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bg_sray, sizeof(float4) * (size_t)R));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bg_wide, sizeof(int4) * (size_t)BG_ITEM_CAP));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->bg_plan, 64));
            size_t tb = 0;
            LRT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int*)nullptr, (int*)nullptr, (int)ncell, s));
            LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_sort_tmp, tb));
            b.ang = (float2*)ctx->bg_ang.p; b.cell_of = (int*)ctx->bg_cell_of.p;
            b.cell_cnt = (int*)ctx->bg_cells.p; b.cell_start = b.cell_cnt + ncell;
            b.sray = (float4*)ctx->bg_sray.p; b.items = (int4*)ctx->bg_wide.p;
            b.plan = (BgPlan*)ctx->bg_plan.p; b.el_bounds = (int*)((char*)ctx->bg_plan.p + 32); b.item_count = b.el_bounds + 2;
            LRT_CUDA_TRY(ctx, cudaMemsetAsync(b.el_bounds, 0x80, sizeof(int) * 2, s));
            LRT_CUDA_TRY(ctx, cudaMemsetAsync(b.item_count, 0, sizeof(int), s));
            LRT_CUDA_TRY(ctx, cudaMemsetAsync(b.cell_cnt, 0, sizeof(int) * ncell, s));
            ctx->span_begin("k_bg_grid", s);
            k_bg_angles<<<(R + 255) / 256, 256, 0, s>>>(a, w, b);
            k_bg_plan<<<1, 32, 0, s>>>(R, b, 0.01f * (float)ctx->opt_beam_cell_pct);
            k_bg_count<<<(R + 255) / 256, 256, 0, s>>>(R, b);
            LRT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->wf_sort_tmp.p, tb, (const int*)b.cell_cnt, b.cell_start, (int)ncell, s));
            k_bg_fill<<<(R + 255) / 256, 256, 0, s>>>(a, b);
            ctx->span_end(s);
            ctx->span_begin("k_bg_bin", s); k_bg_bin<<<G, 256, 0, s>>>(bv, a, w, b, ctx->P_pad); ctx->span_end(s);
            ctx->span_begin("k_bg_heavy", s); k_bg_heavy<<<G, 256, 0, s>>>(bv, a, w, b); ctx->span_end(s);
            ctx->launches += 8;
        } else {
        ctx->span_begin("k_wf_setup", s); k_wf_setup<<<(R + 255) / 256, 256, 0, s>>>(a, w); ctx->span_end(s);
        const uint2* in = nullptr; const int* in_count = nullptr;
        uint2* bufs[2] = {w.list_a, w.list_b};
        int flip = 0;
        for (int level = bv.levels - 1; level >= 1; level--) {
            ctx->span_begin("k_wf_level", s); k_wf_level<<<G, 256, 0, s>>>(bv, a, w, level, in, in_count, bufs[flip], w.counts + (level - 1)); ctx->span_end(s);
            in = bufs[flip]; in_count = w.counts + (level - 1); flip ^= 1;
        }
        ctx->span_begin("k_wf_leaf", s); k_wf_leaf<<<G, 256, 0, s>>>(bv, a, w, in, in_count); ctx->span_end(s);
        ctx->launches += 1 + bv.levels;
        }
        if (ctx->opt_wavefront_shade == 0) {
            ctx->span_begin("k_wf_shade", s); k_wf_shade<<<min((S + 3) / 4, ctx->num_sms * 8), 128, 0, s>>>(bv, a, w); ctx->span_end(s);
        } else {
            if (ctx->opt_sort_rays && ctx->opt_wavefront_shade != 3) {
                // rays by descending candidate count (14-bit keys: 2 radix passes over R pairs)
                int* order = (int*)ctx->wf_ids.p + R;
                size_t tb = 0;
                LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, (const int*)w.hit_count, (int*)ctx->wf_keys.p,
                                                                            (const int*)w.ray_ids, order, R, 0, 14, s));
                LRT_CUDA_TRY(ctx, ctx->reserve(ctx->wf_sort_tmp, tb));
                ctx->span_begin("ray_order_sort", s);
                LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairsDescending(ctx->wf_sort_tmp.p, tb, (const int*)w.hit_count, (int*)ctx->wf_keys.p,
                                                                            (const int*)w.ray_ids, order, R, 0, 14, s));
                ctx->span_end(s);
                w.order = order;
            }
            if (ctx->opt_wavefront_shade == 3) {
                // ---- split passes (lrt_split.cuh): sort + gather into one sorted record stream, slots / colour / fold
                SpBufs sp;
                long long capacity = (long long)R * 96; if (capacity < (1LL << 20)) capacity = 1LL << 20;
                sp.capacity = capacity;
                LRT_CUDA_TRY(ctx, ctx->reserve(ctx->sp_cnt, sizeof(int) * 2 * ((size_t)R + 1)));
                LRT_CUDA_TRY(ctx, ctx->reserve(ctx->sp_rec, sizeof(float4) * 4 * (size_t)capacity));
                sp.ccnt = (int*)ctx->sp_cnt.p; sp.cbase = sp.ccnt + (R + 1); sp.srec = (float4*)ctx->sp_rec.p;
                sp.tri = ctx->opt_triangle_depth; sp.mod = mod; sp.means = means; sp.scales = scales; sp.rots = rots; sp.opac = opac;
                if (sp.tri && (!means || !scales || !rots || !opac)) { ctx->set_error("lrt_forward: LRT_OPT_TRIANGLE_DEPTH needs the Gaussian parameters"); return LRT_ERR_INVALID; }
                size_t tb = 0;
                LRT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int*)sp.ccnt, sp.cbase, R + 1, s));
                LRT_CUDA_TRY(ctx, ctx->reserve(ctx->sp_scan_tmp, tb));
                // the passes talk through the hit lists: the caller's (kept for the backward) or the context's own
                if (!a.hit_gidx || !a.hit_aux) {
                    const int icap = a.hit_gidx ? a.cap : LRT_INTERNAL_HIT_CAP;
                    const size_t per = (size_t)icap * R;
                    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->sp_hits, per * (sizeof(int32_t) + sizeof(float) + sizeof(float4)) + sizeof(int32_t) * (size_t)R + 64));
                    char* base_ = (char*)ctx->sp_hits.p;
                    float4* i_aux = (float4*)base_; int32_t* i_g = (int32_t*)(base_ + per * sizeof(float4));
                    float* i_t = (float*)(i_g + per); int32_t* i_c = (int32_t*)(i_t + per);
                    if (!a.hit_gidx) { a.hit_gidx = i_g; a.hit_t = i_t; a.hit_cnt = i_c; a.cap = icap; }
                    a.hit_aux = i_aux;
                }
                const bool fused = ctx->opt_split_fused && !sp.tri;
                if (fused) {
                    // one warp per ray: sort + the rounds' slot logic, no record stream (lrt_split.cuh)
                    ctx->span_begin("k_sp_warp", s);
                    k_sp_warp<<<min((R + 3) / 4, ctx->num_sms * 24), 128, 0, s>>>(bv, a, w);
                    {   // bins beyond 512 candidates, and the rays with candidates on the overflow list: one block sorts, one warp walks
                        const int smem_keys = 2 * WF_HCAP_MAX;
                        const size_t smem = sizeof(unsigned long long) * (size_t)smem_keys;
                        LRT_CUDA_TRY(ctx, cudaFuncSetAttribute(k_wf_sort_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        k_ov_scatter<<<ctx->num_sms, 256, 0, s>>>(w);
                        k_wf_sort_big<<<ctx->num_sms, 256, smem, s>>>(a, w, smem_keys);
                        k_sp_big<<<ctx->num_sms, 128, 0, s>>>(bv, a, w);
                    }
                    ctx->span_end(s);
                } else {
                    ctx->span_begin("k_sp_sort", s);
                    k_sp_counts<<<(R + 1 + 255) / 256, 256, 0, s>>>(R, w, sp);
                    LRT_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->sp_scan_tmp.p, tb, (const int*)sp.ccnt, sp.cbase, R + 1, s));
                    k_sp_sort<<<min((R + 3) / 4, ctx->num_sms * 16), 128, 0, s>>>(bv, a, w, sp);
                    {   // the few bins beyond 512 candidates (and rays with an overflow list): one block each sorts, one warp walks
                        const int smem_keys = 2 * WF_HCAP_MAX;
                        const size_t smem = sizeof(unsigned long long) * (size_t)smem_keys;
                        LRT_CUDA_TRY(ctx, cudaFuncSetAttribute(k_wf_sort_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        k_ov_scatter<<<ctx->num_sms, 256, 0, s>>>(w);
                        k_wf_sort_big<<<ctx->num_sms, 256, smem, s>>>(a, w, smem_keys);
                        k_sp_big<<<ctx->num_sms, 128, 0, s>>>(bv, a, w);
                    }
                    ctx->span_end(s);
                    ctx->span_begin("k_sp_slots", s);
                    if (sp.tri) {
                        LRT_CUDA_TRY(ctx, cudaFuncSetAttribute(k_sp_slots<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SLOTS_SMEM));
                        k_sp_slots<true><<<(R + 127) / 128, 128, SP_SLOTS_SMEM, s>>>(a, w, sp);
                    } else {
                        LRT_CUDA_TRY(ctx, cudaFuncSetAttribute(k_sp_slots<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SLOTS_SMEM));
                        k_sp_slots<false><<<(R + 127) / 128, 128, SP_SLOTS_SMEM, s>>>(a, w, sp);
                    }
                    ctx->span_end(s);
                }
                ctx->span_begin("k_sp_colour", s);
                if (a.sh_tab && ctx->sh_parts_vec) k_sp_colour<2><<<dim3((R + 255) / 256, 16), 256, 0, s>>>(a);
                else if (sh_rows_aligned(a)) k_sp_colour<1><<<dim3((R + 255) / 256, 16), 256, 0, s>>>(a);
                else k_sp_colour<0><<<dim3((R + 255) / 256, 16), 256, 0, s>>>(a);
                ctx->span_end(s);
                ctx->span_begin("k_sp_fold", s); k_sp_fold<<<(R + 127) / 128, 128, 0, s>>>(a); ctx->span_end(s);
                ctx->launches += 10;
            } else {
                ctx->span_begin("k_wf_sort", s);
                k_wf_sort<<<min((R + 3) / 4, ctx->num_sms * 16), 128, 0, s>>>(a, w);
                {   // the few bins beyond 512 candidates: one block each, keys in dynamic shared memory (64 KB at hcap = 8192)
                    const size_t smem = sizeof(unsigned long long) * (size_t)w.hcap;
                    LRT_CUDA_TRY(ctx, cudaFuncSetAttribute(k_wf_sort_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k_wf_sort_big<<<ctx->num_sms, 256, smem, s>>>(a, w, w.hcap);
                }
                ctx->span_end(s);
                ctx->launches += 1;
                ctx->span_begin("k_wf_composite", s);
                if (ctx->opt_wavefront_shade == 2) k_wf_composite2<<<(S + 127) / 128, 128, 0, s>>>(bv, a, w);
                else k_wf_composite<<<(S + 127) / 128, 128, 0, s>>>(bv, a, w);
                ctx->span_end(s);
                ctx->launches += 1;
            }
        }
        ctx->span_begin("k_wf_fallback", s);
        k_wf_fallback<<<ctx->num_sms, 128, 0, s>>>(bv, a, w);
        if (ctx->opt_wavefront_shade == 3) {        // rays whose hit list outgrew cap: redone per ray, weights already accumulated
            FwdArgs a2 = a; a2.accum_w = nullptr;
            k_wf_fallback_overflow<<<ctx->num_sms, 128, 0, s>>>(bv, a2, w);
            ctx->launches += 1;
        }
        ctx->span_end(s);
        ctx->launches += 2;
    } else if (ctx->opt_forward_kernel == 2) {
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->counter, sizeof(int) * 4));
        LRT_CUDA_TRY(ctx, cudaMemsetAsync(ctx->counter.p, 0, sizeof(int) * 4, s));
        a.work_counter = (int*)ctx->counter.p;
        if (ctx->g8_blocks_per_sm == 0) {
            int nb = 0, sms = 0;
            LRT_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_forward_g8, TB, 0));
            LRT_CUDA_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
            ctx->g8_blocks_per_sm = nb > 0 ? nb : 1; ctx->num_sms = sms > 0 ? sms : 148;
        }
        const int grid = min(ctx->num_sms * ctx->g8_blocks_per_sm, (S + 15) / 16);      // 16 rays in flight per 128-thread block
        ctx->span_begin("k_forward_g8", s); k_forward_g8<<<grid, TB, 0, s>>>(ctx->view(), a); ctx->span_end(s);
    } else {
        LRT_CUDA_TRY(ctx, ctx->reserve(ctx->counter, sizeof(int) * 4));
        LRT_CUDA_TRY(ctx, cudaMemsetAsync(ctx->counter.p, 0, sizeof(int) * 4, s));
        a.work_counter = (int*)ctx->counter.p;
        if (ctx->fwd_blocks_per_sm == 0) {
            int nb = 0, sms = 0;
            LRT_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_forward_persistent, TB, 0));
            LRT_CUDA_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
            ctx->fwd_blocks_per_sm = nb > 0 ? nb : 1; ctx->num_sms = sms > 0 ? sms : 148;
        }
        const int grid = min(ctx->num_sms * ctx->fwd_blocks_per_sm, (S + TB - 1) / TB);   // one resident wave, sized to the SM count
        ctx->span_begin("k_forward_persistent", s); k_forward_persistent<<<grid, TB, 0, s>>>(ctx->view(), a); ctx->span_end(s);
    }
    ctx->launches += 1;
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}
