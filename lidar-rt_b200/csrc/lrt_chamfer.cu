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
// How: both clouds are sorted along a 48-bit Morton curve (16 bits per axis on cubic cells) and get an implicit 8-wide
// hierarchy like the tracer's (level 0 = runs of 8 sorted points, level l node j = union of nodes 8j..8j+7 below).
// A query walks the other cloud's hierarchy depth-first, nearest child first, pruning with the box distance computed
// with the SAME fp32 expression as d — every rounding in it is monotone, so box_distance <= d for every point inside and
// a subtree is skipped only if box_distance > best (ties are still visited: the lowest index must win).  Queries run
// in their own cloud's Morton order, so the 32 lanes of a warp walk nearly the same nodes.
// ~10^2 distance evaluations per point instead of m = 1.5·10^5.
#include <cub/cub.cuh>
#include <cfloat>
#include "lrt_ctx.cuh"

namespace {

constexpr int CH_TB = 256;
constexpr int CH_KEY_BITS = 48;

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

// blockIdx.y selects the cloud; b[6*set + 0..2] = min, 3..5 = max (order-preserving int encoding)
__global__ void __launch_bounds__(CH_TB) k_ch_bounds(int n0, const float* __restrict__ p0, int n1, const float* __restrict__ p1, int* __restrict__ b)
{
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
        for (int a = 0; a < 3; a++) { atomicMin(&b[6 * set + a], f2ord(lo[a])); atomicMax(&b[6 * set + 3 + a], f2ord(hi[a])); }
    }
}

__device__ __forceinline__ unsigned long long spread16(unsigned v)
{
    unsigned long long x = v & 0xffffu;
    x = (x | (x << 16)) & 0x0000ff0000ffull;
    x = (x | (x << 8)) & 0x00f00f00f00full;
    x = (x | (x << 4)) & 0x0c30c30c30c3ull;
    x = (x | (x << 2)) & 0x249249249249ull;
    return x;
}

__global__ void __launch_bounds__(CH_TB) k_ch_keys(int n, const float* __restrict__ p, const int* __restrict__ b,
                                                   unsigned long long* __restrict__ keys, unsigned* __restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lo[3] = {ord2f(b[0]), ord2f(b[1]), ord2f(b[2])};
    const float ext = fmaxf(fmaxf(ord2f(b[3]) - lo[0], ord2f(b[4]) - lo[1]), fmaxf(ord2f(b[5]) - lo[2], 1e-30f));
    const float inv = 65536.0f / ext;
    unsigned q[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float t = fminf(fmaxf((p[3 * (size_t)i + a] - lo[a]) * inv, 0.0f), 65535.0f);     // NaN -> 0
        q[a] = (unsigned)t;
    }
    keys[i] = spread16(q[0]) | (spread16(q[1]) << 1) | (spread16(q[2]) << 2);
    idx[i] = (unsigned)i;
}

// one thread per run of 8 sorted points: gathers them, writes the padded point array and the run's box
__global__ void __launch_bounds__(CH_TB) k_ch_leaves(int n, int n_leaf_pad, const unsigned* __restrict__ order, const float* __restrict__ p,
                                                     float4* __restrict__ pts, float4* __restrict__ boxes)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_leaf_pad) return;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (8 * (long long)j < n) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int s = 8 * j + c;
            float4 v = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(INT_MAX));
            if (s < n) {
                const unsigned g = order[s];
                v = make_float4(p[3 * (size_t)g], p[3 * (size_t)g + 1], p[3 * (size_t)g + 2], __int_as_float((int)g));
                lo[0] = fminf(lo[0], v.x); lo[1] = fminf(lo[1], v.y); lo[2] = fminf(lo[2], v.z);
                hi[0] = fmaxf(hi[0], v.x); hi[1] = fmaxf(hi[1], v.y); hi[2] = fmaxf(hi[2], v.z);
            }
            pts[s] = v;
        }
    }
    boxes[2 * (size_t)j] = make_float4(lo[0], lo[1], lo[2], 0.f);
    boxes[2 * (size_t)j + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
}

// one thread per node of a level above the leaves: union of its 8 children (empty boxes are (+inf, -inf) and drop out)
__global__ void __launch_bounds__(CH_TB) k_ch_fit(int n_child, int n_node_pad, const float4* __restrict__ child, float4* __restrict__ node)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_node_pad) return;
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

// d of the reference build (chamfer3D.cu:27-30 as compiled: FMUL on y, FFMA on x, FFMA on z); explicit intrinsics so
// that neither -fmad setting changes it
__device__ __forceinline__ float ch_dist(float x, float y, float z) { return __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y))); }

// One thread per query point, taken in the query cloud's own Morton order.
__global__ void __launch_bounds__(CH_TB) k_ch_query(int nq, const float4* __restrict__ qpts, ChView T, float* __restrict__ dist, int* __restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const float4 q = __ldg(&qpts[i]);
    float best = INFINITY;
    int bi = INT_MAX;
    unsigned long long stk[8 * LRT_CH_MAX_LEVELS];      // (box distance bits << 32) | level << 28 | node
    int sp = 0;
    int level = T.levels - 1, node = 0;                 // the root; its own box is not tested
    while (true) {
        if (level == 0) {
            const float4* __restrict__ run = T.pts + 8 * (size_t)node;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float4 p = __ldg(&run[c]);
                const float d = ch_dist(__fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y), __fsub_rn(p.z, q.z));
                const int g = __float_as_int(p.w);
                if (d < best || (d == best && g < bi)) { best = d; bi = g; }
            }
        } else {
            const float4* __restrict__ cb = T.boxes + 2 * ((size_t)T.level_off[level - 1] + 8 * (size_t)node);
            float bd[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float4 lo = __ldg(&cb[2 * c]), hi = __ldg(&cb[2 * c + 1]);
                const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, q.x), __fsub_rn(q.x, hi.x)), 0.f);
                const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, q.y), __fsub_rn(q.y, hi.y)), 0.f);
                const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, q.z), __fsub_rn(q.z, hi.z)), 0.f);
                bd[c] = ch_dist(dx, dy, dz);                 // empty boxes: +inf
            }
            int near = 0;
#pragma unroll
            for (int c = 1; c < 8; c++) if (bd[c] < bd[near]) near = c;
            const float bn = bd[near];
            if (bn <= best && bn < INFINITY) {
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (c != near && bd[c] <= best && bd[c] < INFINITY)
                        stk[sp++] = ((unsigned long long)__float_as_uint(bd[c]) << 32) | ((unsigned)(level - 1) << 28) | (unsigned)(8 * node + c);
                }
                level -= 1; node = 8 * node + near;
                continue;
            }
        }
        bool found = false;
        while (sp > 0) {
            const unsigned long long e = stk[--sp];
            if (__uint_as_float((unsigned)(e >> 32)) <= best) {
                level = (int)(((unsigned)e) >> 28); node = (int)(((unsigned)e) & 0x0fffffffu);
                found = true;
                break;
            }
        }
        if (!found) break;
    }
    const int o = __float_as_int(q.w);
    dist[o] = best;
    idx[o] = bi == INT_MAX ? 0 : bi;
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

int ch_build_tree(lrt_ctx* ctx, lrt_ctx::ChTree& t, int n, const float* xyz, const int* bounds, cudaStream_t s)
{
    t.n = n;
    t.n_pad = (n + 7) / 8 * 8;
    int cnt = t.n_pad / 8, off = 0, L = 0;
    while (true) {
        if (L >= LRT_CH_MAX_LEVELS) { ctx->set_error("lrt_chamfer: too many points"); return LRT_ERR_INVALID; }
        t.level_cnt[L] = cnt; t.level_off[L] = off;
        const int pad = cnt == 1 ? 1 : (cnt + 7) / 8 * 8;
        off += pad; L++;
        if (cnt == 1) break;
        cnt = (cnt + 7) / 8;
    }
    t.levels = L;
    LRT_CUDA_TRY(ctx, ctx->reserve(t.keys_a, sizeof(unsigned long long) * (size_t)n));
    LRT_CUDA_TRY(ctx, ctx->reserve(t.keys_b, sizeof(unsigned long long) * (size_t)n));
    LRT_CUDA_TRY(ctx, ctx->reserve(t.idx_a, sizeof(unsigned) * (size_t)n));
    LRT_CUDA_TRY(ctx, ctx->reserve(t.idx_b, sizeof(unsigned) * (size_t)n));
    LRT_CUDA_TRY(ctx, ctx->reserve(t.pts, sizeof(float4) * (size_t)t.n_pad));
    LRT_CUDA_TRY(ctx, ctx->reserve(t.boxes, sizeof(float4) * 2 * (size_t)off));
    cub::DoubleBuffer<unsigned long long> dk((unsigned long long*)t.keys_a.p, (unsigned long long*)t.keys_b.p);
    cub::DoubleBuffer<unsigned> dv((unsigned*)t.idx_a.p, (unsigned*)t.idx_b.p);
    size_t tmp = 0;
    LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, n, 0, CH_KEY_BITS, s));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->ch_tmp, tmp));
    ctx->span_begin("k_ch_keys", s);
    k_ch_keys<<<(n + CH_TB - 1) / CH_TB, CH_TB, 0, s>>>(n, xyz, bounds, dk.Current(), dv.Current());
    ctx->span_end(s);
    ctx->span_begin("ch_radix_sort", s);
    LRT_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->ch_tmp.p, tmp, dk, dv, n, 0, CH_KEY_BITS, s));
    ctx->span_end(s);
    float4* boxes = (float4*)t.boxes.p;
    const int leaf_pad = t.levels == 1 ? 1 : (t.level_cnt[0] + 7) / 8 * 8;
    ctx->span_begin("k_ch_leaves", s);
    k_ch_leaves<<<(leaf_pad + CH_TB - 1) / CH_TB, CH_TB, 0, s>>>(n, leaf_pad, dv.Current(), xyz, (float4*)t.pts.p, boxes);
    ctx->span_end(s);
    ctx->span_begin("k_ch_fit", s);
    for (int l = 1; l < t.levels; l++) {
        const int pad = t.level_cnt[l] == 1 ? 1 : (t.level_cnt[l] + 7) / 8 * 8;
        k_ch_fit<<<(pad + CH_TB - 1) / CH_TB, CH_TB, 0, s>>>(t.level_cnt[l - 1], pad, boxes + 2 * (size_t)t.level_off[l - 1], boxes + 2 * (size_t)t.level_off[l]);
    }
    ctx->span_end(s);
    ctx->launches += 1 + 2 + CH_KEY_BITS / 8 + 1 + (t.levels - 1);     // keys, radix sort (histogram, scan, one pass per digit), leaves, fit
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
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->ch_bounds, sizeof(int) * 12));
    int* bounds = (int*)ctx->ch_bounds.p;
    for (int i = 0; i < b; i++) {
        const float* a = xyz1 + 3 * (size_t)i * n;
        const float* c = xyz2 + 3 * (size_t)i * m;
        ctx->span_begin("k_ch_bounds", s);
        k_ch_bounds_init<<<1, 32, 0, s>>>(bounds);
        const int nb = std::min(std::max((std::max(n, m) + CH_TB - 1) / CH_TB, 1), 4 * 148);
        k_ch_bounds<<<dim3(nb, 2), CH_TB, 0, s>>>(n, a, m, c, bounds);
        ctx->span_end(s);
        ctx->launches += 2;
        int rc = ch_build_tree(ctx, ctx->ch[0], n, a, bounds, s);
        if (rc != LRT_OK) return rc;
        rc = ch_build_tree(ctx, ctx->ch[1], m, c, bounds + 6, s);
        if (rc != LRT_OK) return rc;
        ctx->span_begin("k_ch_query", s);
        k_ch_query<<<(n + CH_TB - 1) / CH_TB, CH_TB, 0, s>>>(n, (const float4*)ctx->ch[0].pts.p, ch_view(ctx->ch[1]), dist1 + (size_t)i * n, idx1 + (size_t)i * n);
        k_ch_query<<<(m + CH_TB - 1) / CH_TB, CH_TB, 0, s>>>(m, (const float4*)ctx->ch[1].pts.p, ch_view(ctx->ch[0]), dist2 + (size_t)i * m, idx2 + (size_t)i * m);
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
