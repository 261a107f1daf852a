// lrt_chamfer.cu — Chamfer distance (nearest neighbour in both directions) between two point clouds, and its VJP.
//
// SURVEY.md §8(f) N2. Replaces lib/utils/chamfer3D/chamfer3D.cu of the reference: NmDistanceKernel (:11-133, an
// O(n·m) scan of all of xyz2 for every point of xyz1 through 512-point shared-memory batches, launched once per
// direction :144-145) and NmDistanceGradKernel (:157-178, six float atomics per point, launched once per direction).
//
// What must come out (and does, bit for bit, see tests/test_chamfer.py):
//   dist[j] = min_k d(j,k),  d = fma(z,z, fma(x,x, y*y)) with (x,y,z) = xyz2[k] - xyz1[j]
//             (the contraction nvcc chooses for `x2*x2+y2*y2+z2*z2` in the reference build: FMUL y, FFMA x, FFMA z)
//   idx[j]  = the LOWEST k attaining it (the reference compares with strict `<` inside a batch, :31-33, and strict `>`
//             across batches, :121, so the first occurrence wins).
//
// How: both clouds are sorted along a 39-bit Morton curve (13 bits per axis on cubic cells; one radix sort orders both) and
// get an implicit 8-wide hierarchy like the tracer's (level 0 = runs of 8 sorted points, level l node j = union of nodes
// 8j..8j+7 below).
// A query walks the other cloud's hierarchy depth-first, nearest child first, pruning with the box distance computed
// with the SAME fp32 expression as d — every rounding in it is monotone, so box_distance <= d for every point inside and
// a subtree is skipped only if box_distance > best (ties are still visited: the lowest index must win).  Queries run
// in their own cloud's Morton order, eight lanes per query, so the four queries of a warp walk nearly the same nodes.
// ~10^2 distance evaluations per point instead of m = 1.5·10^5.
#include <cub/cub.cuh>
#include <cfloat>
#include "lrt_ctx.cuh"

namespace {

constexpr int CH_TB = 256;
constexpr int CH_KEY_BITS = 40;       // 13 bits per axis + the cloud bit

struct ChView {
    const float4* pts;      // sorted points (x, y, z, original index as int bits), padded to a multiple of 8 with +inf
    const float4* boxes;    // two float4 per node: (lo.xyz, -) (hi.xyz, -); levels concatenated, each padded to a multiple of 8
    int level_off[LRT_CH_MAX_LEVELS];
    int levels;             // number of box levels; the root is the single node of level levels-1
    int n;
};

__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_ch_bounds_init(int* b)
{
    const int t = threadIdx.x;
    if (t < 12) b[t] = (t % 6) < 3 ? INT_MAX : INT_MIN;
}

// blockIdx.y selects the cloud; b[6*set + 0..2] = min, 3..5 = max (order-preserving int encoding); one set of atomics per block
__global__ void __launch_bounds__(CH_TB) k_ch_bounds(int n0, const float* __restrict__ p0, int n1, const float* __restrict__ p1, int* __restrict__ b)
{
    __shared__ float red[CH_TB / 32][6];
    const int set = blockIdx.y;
    const int n = set ? n1 : n0;
    const float* __restrict__ p = set ? p1 : p0;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; a++) { const float v = p[3 * (size_t)i + a]; lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { red[threadIdx.x >> 5][a] = lo[a]; red[threadIdx.x >> 5][3 + a] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[0][threadIdx.x];
        for (int w = 1; w < CH_TB / 32; w++) v = threadIdx.x < 3 ? fminf(v, red[w][threadIdx.x]) : fmaxf(v, red[w][threadIdx.x]);
        if (threadIdx.x < 3) atomicMin(&b[6 * set + threadIdx.x], f2ord(v)); else atomicMax(&b[6 * set + threadIdx.x], f2ord(v));
    }
}

__device__ __forceinline__ unsigned long long spread13(unsigned v)
{
    unsigned long long x = v & 0x1fffu;
    x = (x | (x << 16)) & 0x0000ff0000ffull;
    x = (x | (x << 8)) & 0x00f00f00f00full;
    x = (x | (x << 4)) & 0x0c30c30c30c3ull;
    x = (x | (x << 2)) & 0x249249249249ull;
    return x;
}

// keys of both clouds in one array: (cloud << 39) | 39-bit Morton code in the cloud's own cubic-cell grid; value = position in
// the concatenation (cloud 1 starts at n0). One radix sort then orders both clouds.
__global__ void __launch_bounds__(CH_TB) k_ch_keys(int n0, const float* __restrict__ p0, int n1, const float* __restrict__ p1, const int* __restrict__ b,
                                                   unsigned long long* __restrict__ keys, unsigned* __restrict__ idx)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n0 + n1) return;
    const int set = t >= n0;
    const int i = set ? t - n0 : t;
    const float* __restrict__ p = set ? p1 : p0;
    b += 6 * set;
    const float lo[3] = {ord2f(b[0]), ord2f(b[1]), ord2f(b[2])};
    const float ext = fmaxf(fmaxf(ord2f(b[3]) - lo[0], ord2f(b[4]) - lo[1]), fmaxf(ord2f(b[5]) - lo[2], 1e-30f));
    const float inv = 8192.0f / ext;
    unsigned q[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float v = fminf(fmaxf((p[3 * (size_t)i + a] - lo[a]) * inv, 0.0f), 8191.0f);     // NaN -> 0
        q[a] = (unsigned)v;
    }
    keys[t] = ((unsigned long long)set << 39) | spread13(q[0]) | (spread13(q[1]) << 1) | (spread13(q[2]) << 2);
    idx[t] = (unsigned)t;
}

struct ChBuild {
    int n, base;            // points of the cloud; its offset in the sorted concatenation
    const float* xyz;
    float4* pts;
    float4* boxes;
    int level_off[LRT_CH_MAX_LEVELS], level_cnt[LRT_CH_MAX_LEVELS], levels;
};

__device__ __forceinline__ int ch_pad(int cnt) { return cnt == 1 ? 1 : (cnt + 7) / 8 * 8; }

// one thread per run of 8 sorted points (blockIdx.y = cloud): gathers them, writes the padded point array and the run's box
__global__ void __launch_bounds__(CH_TB) k_ch_leaves(ChBuild t0, ChBuild t1, const unsigned* __restrict__ order)
{
    const ChBuild& t = blockIdx.y ? t1 : t0;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ch_pad(t.level_cnt[0])) return;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (8 * (long long)j < t.n) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int s = 8 * j + c;
            float4 v = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(INT_MAX));
            if (s < t.n) {
                const unsigned g = order[t.base + s] - (unsigned)t.base;
                v = make_float4(t.xyz[3 * (size_t)g], t.xyz[3 * (size_t)g + 1], t.xyz[3 * (size_t)g + 2], __int_as_float((int)g));
                lo[0] = fminf(lo[0], v.x); lo[1] = fminf(lo[1], v.y); lo[2] = fminf(lo[2], v.z);
                hi[0] = fmaxf(hi[0], v.x); hi[1] = fmaxf(hi[1], v.y); hi[2] = fmaxf(hi[2], v.z);
            }
            t.pts[s] = v;
        }
    }
    t.boxes[2 * (size_t)j] = make_float4(lo[0], lo[1], lo[2], 0.f);
    t.boxes[2 * (size_t)j + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
}

// node j of a level above the leaves: union of its 8 children (empty boxes are (+inf, -inf) and drop out)
__device__ __forceinline__ void ch_fit_node(int j, int n_child, const float4* __restrict__ child, float4* __restrict__ node)
{
    float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int k = 8 * j + c;
        if (k < n_child) {
            const float4 a = child[2 * (size_t)k], b = child[2 * (size_t)k + 1];
            lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
            hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
        }
    }
    node[2 * (size_t)j] = lo; node[2 * (size_t)j + 1] = hi;
}

// level 1 of both clouds, one thread per node
__global__ void __launch_bounds__(CH_TB) k_ch_fit1(ChBuild t0, ChBuild t1)
{
    const ChBuild& t = blockIdx.y ? t1 : t0;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (t.levels < 2 || j >= ch_pad(t.level_cnt[1])) return;
    ch_fit_node(j, t.level_cnt[0], t.boxes, t.boxes + 2 * (size_t)t.level_off[1]);
}

// levels 2.. of a cloud in one block (1/64 of the points and shrinking by 8 per level)
__global__ void __launch_bounds__(1024) k_ch_fit_top(ChBuild t0, ChBuild t1)
{
    const ChBuild& t = blockIdx.x ? t1 : t0;
    for (int l = 2; l < t.levels; l++) {
        for (int j = threadIdx.x; j < ch_pad(t.level_cnt[l]); j += blockDim.x)
            ch_fit_node(j, t.level_cnt[l - 1], t.boxes + 2 * (size_t)t.level_off[l - 1], t.boxes + 2 * (size_t)t.level_off[l]);
        __syncthreads();
    }
}

// d of the reference build (chamfer3D.cu:27-30 as compiled: FMUL on y, FFMA on x, FFMA on z); explicit intrinsics so
// that neither -fmad setting changes it
__device__ __forceinline__ float ch_dist(float x, float y, float z) { return __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y))); }

// Eight lanes per query point (four queries per warp), queries taken in their own cloud's Morton order.
// A visit to a node is one step of the group: lane c loads child c — a 32-byte box, or a 16-byte point of a level-0 node,
// treated as the box lo = hi = p so that both take the same arithmetic — the group's minimum is one redux.sync, and the
// group descends into the nearest child whose box distance is <= best, or records the nearest point (lowest index among
// equals: a second redux over the indices of the lanes at the minimum). No stack: the box distances of a node's children
// stay with the lanes (one float per lane and level, in shared memory) until the walk climbs back through that node, so
// a node is loaded once; the parent of node j is j >> 3, and climbing costs one shared-memory load and one redux per level.
// Starting bound: the target point with the query's own index, when there is one. Any target point is a valid upper bound,
// so results do not depend on it; for the clouds this is called with (the same rays back-projected with predicted and
// measured range, train.py:198-205; range2point of two range images, metric_utils.py:450-451) it is already within
// centimetres of the answer and the walk only has to rule the rest out.
__global__ void __launch_bounds__(CH_TB) k_ch_query(int nq, const float4* __restrict__ qpts, ChView T, const float* __restrict__ txyz,
                                                    float* __restrict__ dist, int* __restrict__ idx)
{
    __shared__ float pend[LRT_CH_MAX_LEVELS][CH_TB];
    __shared__ int s_off[LRT_CH_MAX_LEVELS];
    if (threadIdx.x < LRT_CH_MAX_LEVELS) s_off[threadIdx.x] = T.level_off[threadIdx.x];
    __syncthreads();
    const int lane8 = threadIdx.x & 7;
    const unsigned gmask = 0xffu << (threadIdx.x & 24);
    const long long gq = ((long long)blockIdx.x * CH_TB + threadIdx.x) >> 3;
    if (gq >= nq) return;                                   // whole groups leave together
    const float4 q = __ldg(&qpts[gq]);
    const unsigned none = 0xffffffffu;
    float best = INFINITY;
    int bi = INT_MAX;
    {
        const int o = __float_as_int(q.w);
        if (o < T.n) {
            const float* __restrict__ p = txyz + 3 * (size_t)o;
            best = ch_dist(__fsub_rn(__ldg(p), q.x), __fsub_rn(__ldg(p + 1), q.y), __fsub_rn(__ldg(p + 2), q.z));
            bi = o;
        }
    }
    const int top = T.levels - 1;
    int level = top, node = 0;
    while (true) {
        // ---- arrive at (level, node) from above: lane c evaluates child c
        const bool leaf = level == 0;
        const float4* __restrict__ ptr = leaf ? T.pts + 8 * (size_t)node + lane8 : T.boxes + 2 * ((size_t)s_off[leaf ? 0 : level - 1] + 8 * (size_t)node + lane8);
        const float4 lo = __ldg(ptr);
        float4 hi = lo;
        if (!leaf) hi = __ldg(ptr + 1);
        const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, q.x), __fsub_rn(q.x, hi.x)), 0.f);     // leaf: |p.x - q.x|
        const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, q.y), __fsub_rn(q.y, hi.y)), 0.f);
        const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, q.z), __fsub_rn(q.z, hi.z)), 0.f);
        const float bd = ch_dist(dx, dy, dz);                // empty boxes and padding points: +inf
        unsigned k;
        if (leaf) {
            const unsigned dmin = __reduce_min_sync(gmask, __float_as_uint(bd));          // d >= 0: its bits order like the value
            const int imin = __reduce_min_sync(gmask, __float_as_uint(bd) == dmin ? __float_as_int(lo.w) : INT_MAX);
            const float d = __uint_as_float(dmin);
            if (d < best || (d == best && imin < bi)) { best = d; bi = imin; }
            k = none;
        } else {
            pend[level][threadIdx.x] = bd;
            k = __reduce_min_sync(gmask, (bd <= best && bd < INFINITY) ? ((__float_as_uint(bd) & ~7u) | (unsigned)lane8) : none);
        }
        // ---- nothing (more) to enter here: climb to the nearest ancestor that still has a child worth entering
        while (k == none) {
            level += 1; node >>= 3;
            if (level > top) break;
            const float pd = pend[level][threadIdx.x];
            k = __reduce_min_sync(gmask, (pd <= best && pd < INFINITY) ? ((__float_as_uint(pd) & ~7u) | (unsigned)lane8) : none);
        }
        if (level > top) break;
        const int c = (int)(k & 7u);
        if (lane8 == c) pend[level][threadIdx.x] = INFINITY;
        node = 8 * node + c; level -= 1;
    }
    if (lane8 == 0) {
        const int o = __float_as_int(q.w);
        dist[o] = best;
        idx[o] = bi == INT_MAX ? 0 : bi;
    }
}

// VJP, chamfer3D.cu:157-178.  Pass 1 (plain stores, doubles as the zero-fill the reference's torch.zeros does): the term a
// point receives as the query of its own direction.  Pass 2 (float atomics): the term its nearest neighbours send back.
__global__ void __launch_bounds__(CH_TB) k_ch_grad_direct(int n, const float* __restrict__ p1, const float* __restrict__ p2, const float* __restrict__ g1,
                                                          const int* __restrict__ i1, float* __restrict__ o1,
                                                          int m, const float* __restrict__ g2, const int* __restrict__ i2, float* __restrict__ o2)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    const float *a = p1, *b = p2, *g = g1; const int* ix = i1; float* o = o1;
    if (j >= n) { j -= n; if (j >= m) return; a = p2; b = p1; g = g2; ix = i2; o = o2; }
    const int k = ix[j];
    const float w = g[j] * 2;
#pragma unroll
    for (int c = 0; c < 3; c++) o[3 * (size_t)j + c] = w * (a[3 * (size_t)j + c] - b[3 * (size_t)k + c]);
}

__global__ void __launch_bounds__(CH_TB) k_ch_grad_scatter(int n, const float* __restrict__ p1, const float* __restrict__ p2, const float* __restrict__ g1,
                                                           const int* __restrict__ i1, float* __restrict__ o1,
                                                           int m, const float* __restrict__ g2, const int* __restrict__ i2, float* __restrict__ o2)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    const float *a = p1, *b = p2, *g = g1; const int* ix = i1; float* o = o2;
    if (j >= n) { j -= n; if (j >= m) return; a = p2; b = p1; g = g2; ix = i2; o = o1; }
    const int k = ix[j];
    const float w = g[j] * 2;
#pragma unroll
    for (int c = 0; c < 3; c++) atomicAdd(&o[3 * (size_t)k + c], -(w * (a[3 * (size_t)j + c] - b[3 * (size_t)k + c])));
}

int ch_layout(lrt_ctx* ctx, lrt_ctx::ChTree& t, int n)
{
    t.n = n;
    t.n_pad = (n + 7) / 8 * 8;
    int cnt = t.n_pad / 8, off = 0, L = 0;
    while (true) {
        if (L >= LRT_CH_MAX_LEVELS) { ctx->set_error("lrt_chamfer: too many points"); return LRT_ERR_INVALID; }
        t.level_cnt[L] = cnt; t.level_off[L] = off;
        off += cnt == 1 ? 1 : (cnt + 7) / 8 * 8; L++;
        if (cnt == 1) break;
        cnt = (cnt + 7) / 8;
    }
    t.levels = L;
    LRT_CUDA_TRY(ctx, ctx->reserve(t.pts, sizeof(float4) * (size_t)t.n_pad));
    LRT_CUDA_TRY(ctx, ctx->reserve(t.boxes, sizeof(float4) * 2 * (size_t)off));
    return LRT_OK;
}

ChBuild ch_build_desc(const lrt_ctx::ChTree& t, int base, const float* xyz)
{
    ChBuild d;
    d.n = t.n; d.base = base; d.xyz = xyz; d.pts = (float4*)t.pts.p; d.boxes = (float4*)t.boxes.p;
    for (int i = 0; i < LRT_CH_MAX_LEVELS; i++) { d.level_off[i] = t.level_off[i]; d.level_cnt[i] = t.level_cnt[i]; }
    d.levels = t.levels;
    return d;
}

// Morton order + hierarchy of both clouds: 8 launches + one radix sort whatever the sizes
int ch_build_trees(lrt_ctx* ctx, int n, const float* a, int m, const float* c, cudaStream_t s)
{
    int rc = ch_layout(ctx, ctx->ch[0], n);
    if (rc != LRT_OK) return rc;
    rc = ch_layout(ctx, ctx->ch[1], m);
    if (rc != LRT_OK) return rc;
    const int tot = n + m;
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->ch_bounds, sizeof(int) * 12));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->ch_keys_a, sizeof(unsigned long long) * (size_t)tot));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->ch_keys_b, sizeof(unsigned long long) * (size_t)tot));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->ch_idx_a, sizeof(unsigned) * (size_t)tot));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->ch_idx_b, sizeof(unsigned) * (size_t)tot));
    cub::DoubleBuffer<unsigned long long> dk((unsigned long long*)ctx->ch_keys_a.p, (unsigned long long*)ctx->ch_keys_b.p);
    cub::DoubleBuffer<unsigned> dv((unsigned*)ctx->ch_idx_a.p, (unsigned*)ctx->ch_idx_b.p);
    size_t tmp = 0;
    LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, tot, 0, CH_KEY_BITS, s));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->ch_tmp, tmp));
    int* bounds = (int*)ctx->ch_bounds.p;
    ctx->span_begin("k_ch_bounds", s);
    k_ch_bounds_init<<<1, 32, 0, s>>>(bounds);
    const int nb = std::min(std::max((std::max(n, m) + CH_TB - 1) / CH_TB, 1), 2 * 148);
    k_ch_bounds<<<dim3(nb, 2), CH_TB, 0, s>>>(n, a, m, c, bounds);
    ctx->span_end(s);
    ctx->span_begin("k_ch_keys", s);
    k_ch_keys<<<(tot + CH_TB - 1) / CH_TB, CH_TB, 0, s>>>(n, a, m, c, bounds, dk.Current(), dv.Current());
    ctx->span_end(s);
    ctx->span_begin("ch_radix_sort", s);
    LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->ch_tmp.p, tmp, dk, dv, tot, 0, CH_KEY_BITS, s));
    ctx->span_end(s);
    const ChBuild t0 = ch_build_desc(ctx->ch[0], 0, a), t1 = ch_build_desc(ctx->ch[1], n, c);
    const int leaves = std::max(ctx->ch[0].level_cnt[0], ctx->ch[1].level_cnt[0]);
    const int leaf_pad = (leaves + 7) / 8 * 8;
    ctx->span_begin("k_ch_leaves", s);
    k_ch_leaves<<<dim3((leaf_pad + CH_TB - 1) / CH_TB, 2), CH_TB, 0, s>>>(t0, t1, dv.Current());
    ctx->span_end(s);
    ctx->span_begin("k_ch_fit", s);
    k_ch_fit1<<<dim3((leaf_pad / 8 + 8 + CH_TB - 1) / CH_TB, 2), CH_TB, 0, s>>>(t0, t1);
    k_ch_fit_top<<<2, 1024, 0, s>>>(t0, t1);
    ctx->span_end(s);
    ctx->launches += 2 + 1 + 2 + CH_KEY_BITS / 8 + 1 + 2;     // bounds, keys, radix sort (histogram, scan, one pass per digit), leaves, fit
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}

ChView ch_view(const lrt_ctx::ChTree& t)
{
    ChView v;
    v.pts = (const float4*)t.pts.p; v.boxes = (const float4*)t.boxes.p;
    for (int i = 0; i < LRT_CH_MAX_LEVELS; i++) v.level_off[i] = t.level_off[i];
    v.levels = t.levels; v.n = t.n;
    return v;
}

} // namespace

int lrt_chamfer_forward_impl(lrt_ctx* ctx, int b, int n, const float* xyz1, int m, const float* xyz2,
                             float* dist1, int32_t* idx1, float* dist2, int32_t* idx2, cudaStream_t s)
{
    if (b < 0 || n < 0 || m < 0 || (long long)b * n > 0x7fffffffLL / 4 || (long long)b * m > 0x7fffffffLL / 4) { ctx->set_error("lrt_chamfer_forward: bad sizes"); return LRT_ERR_INVALID; }
    if ((b > 0 && n > 0 && (!xyz1 || !dist1 || !idx1)) || (b > 0 && m > 0 && (!xyz2 || !dist2 || !idx2))) { ctx->set_error("lrt_chamfer_forward: null argument"); return LRT_ERR_INVALID; }
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (b == 0 || (n == 0 && m == 0)) return LRT_OK;
    if (n == 0 || m == 0) {
        // the reference's kernel never runs its batch loop against an empty cloud: the zero-initialised outputs stay (dist_chamfer_3D.py:43-47)
        if (n) { LRT_CUDA_TRY(ctx, cudaMemsetAsync(dist1, 0, sizeof(float) * (size_t)b * n, s)); LRT_CUDA_TRY(ctx, cudaMemsetAsync(idx1, 0, sizeof(int32_t) * (size_t)b * n, s)); }
        if (m) { LRT_CUDA_TRY(ctx, cudaMemsetAsync(dist2, 0, sizeof(float) * (size_t)b * m, s)); LRT_CUDA_TRY(ctx, cudaMemsetAsync(idx2, 0, sizeof(int32_t) * (size_t)b * m, s)); }
        return LRT_OK;
    }
    for (int i = 0; i < b; i++) {
        const int rc = ch_build_trees(ctx, n, xyz1 + 3 * (size_t)i * n, m, xyz2 + 3 * (size_t)i * m, s);
        if (rc != LRT_OK) return rc;
        ctx->span_begin("k_ch_query", s);
        k_ch_query<<<(int)((8LL * n + CH_TB - 1) / CH_TB), CH_TB, 0, s>>>(n, (const float4*)ctx->ch[0].pts.p, ch_view(ctx->ch[1]), xyz2 + 3 * (size_t)i * m, dist1 + (size_t)i * n, idx1 + (size_t)i * n);
        k_ch_query<<<(int)((8LL * m + CH_TB - 1) / CH_TB), CH_TB, 0, s>>>(m, (const float4*)ctx->ch[1].pts.p, ch_view(ctx->ch[0]), xyz1 + 3 * (size_t)i * n, dist2 + (size_t)i * m, idx2 + (size_t)i * m);
        ctx->span_end(s);
        ctx->launches += 2;
        LRT_CUDA_TRY(ctx, cudaGetLastError());
    }
    return LRT_OK;
}

int lrt_chamfer_backward_impl(lrt_ctx* ctx, int b, int n, const float* xyz1, int m, const float* xyz2,
                              const float* grad_dist1, const float* grad_dist2, const int32_t* idx1, const int32_t* idx2,
                              float* grad_xyz1, float* grad_xyz2, cudaStream_t s)
{
    if (b < 0 || n < 0 || m < 0 || (long long)b * n > 0x7fffffffLL / 4 || (long long)b * m > 0x7fffffffLL / 4) { ctx->set_error("lrt_chamfer_backward: bad sizes"); return LRT_ERR_INVALID; }
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (b == 0 || (n == 0 && m == 0)) return LRT_OK;
    if ((n > 0 && (!xyz1 || !grad_dist1 || !idx1 || !grad_xyz1)) || (m > 0 && (!xyz2 || !grad_dist2 || !idx2 || !grad_xyz2))) { ctx->set_error("lrt_chamfer_backward: null argument"); return LRT_ERR_INVALID; }
    if (n == 0 || m == 0) {      // no pairs: the reference's zero-filled gradients
        if (n) LRT_CUDA_TRY(ctx, cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * 3 * (size_t)b * n, s));
        if (m) LRT_CUDA_TRY(ctx, cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * 3 * (size_t)b * m, s));
        return LRT_OK;
    }
    for (int i = 0; i < b; i++) {
        const size_t on = (size_t)i * n, om = (size_t)i * m;
        const int g = (n + m + CH_TB - 1) / CH_TB;
        ctx->span_begin("k_ch_grad", s);
        k_ch_grad_direct<<<g, CH_TB, 0, s>>>(n, xyz1 + 3 * on, xyz2 + 3 * om, grad_dist1 + on, idx1 + on, grad_xyz1 + 3 * on, m, grad_dist2 + om, idx2 + om, grad_xyz2 + 3 * om);
        k_ch_grad_scatter<<<g, CH_TB, 0, s>>>(n, xyz1 + 3 * on, xyz2 + 3 * om, grad_dist1 + on, idx1 + on, grad_xyz1 + 3 * on, m, grad_dist2 + om, idx2 + om, grad_xyz2 + 3 * om);
        ctx->span_end(s);
        ctx->launches += 2;
    }
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}
