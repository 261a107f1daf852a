// lrt_split.cuh — forward compositing as SPLIT PASSES (default, LRT_OPT_WAVEFRONT_SHADE = 3).
// Included by lrt_forward.cu inside its anonymous namespace, after lrt_wavefront.cuh (shares WfBufs, the candidate bins, the
// register sorts, k_wf_sort_big and the fallbacks).
//
// k_wf_composite2 (one thread per ray doing everything) was bound by exposed load latency: 168 registers, 17 % of the warps
// resident, each thread walking a chain of dependent gathers bin key -> 64 B surfel record -> 192 B SH row (ncu:
// profiles/r1_j_composite_stalls.txt, profiles/r2_a_fwd_chain_before_ncu.txt). The reference's round loop (forward.cu:195-292) is
// only sequential in WHICH hits become slots and how the transmittance evolves; the colour of a hit depends on nothing but the
// ray direction and the Gaussian's SH row. So the work is cut where the dependencies are:
//
// Rays are taken in their natural order by every pass (the by-length ordering of the one-kernel form scattered the hit-list
// writes and the per-Gaussian atomics of neighbouring lanes over unrelated rays: 0.34 -> 0.29 ms for pass A without it).
//
//   k_sp_sort    one WARP per ray (as k_wf_sort): sorts the ray's candidate bin by (t from the origin, id) in registers, then
//                every lane GATHERS the 64-byte surfel record of its candidate and writes it, in sorted order, to the ray's
//                slice of one contiguous stream (slice offsets = exclusive scan of the candidate counts). 32 gathers in flight
//                per warp instead of one per thread; the candidate's t rides in the record's spare word.
//   k_sp_slots   pass A, one THREAD per ray: the round loop — exact re-test from the re-based origin, 16-slot k-buffer,
//                opacity, termination — reading its candidates as a sequential stream (addresses known up front, no key ->
//                record dependency, neighbouring candidates share cache lines). No SH: writes (id, depth, alpha) of every
//                contributing hit, the per-Gaussian weights, and the channels that do not depend on colour (depth, accum, T).
//   k_sp_colour  pass B, one thread per (ray, hit): SH colour of every recorded hit at full occupancy, in natural ray order
//                (neighbouring rays hit the same Gaussians: each SH row comes from DRAM about once per frame instead of once per
//                ray that touches it out of a by-length-sorted order).
//   k_sp_fold    pass C, one thread per ray: the ordered fold C += (alpha T) c over the ray's recorded (alpha, colour), coalesced
//                across rays -> out[0:3].
// The arithmetic of every output is operation for operation that of fwd_shade_round(): results are bit-identical to the other
// forward kernels (tests/test_gpu_parity.py::test_all_forward_kernels_and_options_agree_bitwise).
#pragma once

struct SpBufs {
    int* ccnt;                      // (R + 1) candidates per ray that go through the stream (0: ray handed to the fallback)
    int* cbase;                     // (R + 1) exclusive scan of ccnt = first record of each ray's slice
    float4* srec;                   // the stream: 4 x float4 per candidate = SurfelRec with r3.w = t from the ray's origin
    long long capacity;             // 64-byte records the stream can hold
    // LRT_OPT_TRIANGLE_DEPTH: 8 x float4 per candidate — the record, then the proxy quad's four corners as build2DRectangle
    // rounds them to fp32 (derived again from the caller's raw parameters)
    int tri; float mod;
    const float* means; const float* scales; const float* rots; const float* opac;
};

__global__ void __launch_bounds__(256) k_sp_counts(int R, WfBufs w, SpBufs sp)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > R) return;
    int c = 0;
    if (r < R) { const int hc = w.hit_count[r]; c = ((hc & WF_TAINT) || hc > WF_HCAP) ? 0 : hc; }      // longer bins: fallback or k_sp_big, no stream slice
    sp.ccnt[r] = c;
}

// gather the record of candidate `key` and drop it at sorted position i of the ray's slice of the stream
__device__ __forceinline__ void sp_emit(const SurfelRec* __restrict__ rec_g, const SpBufs& sp, float4* __restrict__ slice, int i, unsigned long long key)
{
    const int g = (int)(unsigned)(key & 0xffffffffull);
    const float4 r0 = ld_f4(&rec_g[g].r0), r1 = ld_f4(&rec_g[g].r1), r2 = ld_f4(&rec_g[g].r2);
    float4 r3 = ld_f4(&rec_g[g].r3);
    r3.w = __uint_as_float((unsigned)(key >> 32));
    float4* dst = slice + (sp.tri ? 8 : 4) * (size_t)i;
    dst[0] = r0; dst[1] = r1; dst[2] = r2; dst[3] = r3;
    if (sp.tri) {
        // the quad's corners exactly as build2DRectangle leaves them in fp32 (primitive_utils.py:203-209; oracle derive()):
        // world = R diag(sx f, sy f, 1) local + mu with local = (-1,1) (-1,-1) (1,1) (1,-1); scale_modifier plays no part
        const float mu[3] = {sp.means[3 * (size_t)g], sp.means[3 * (size_t)g + 1], sp.means[3 * (size_t)g + 2]};
        const float sc[2] = {sp.scales[2 * (size_t)g], sp.scales[2 * (size_t)g + 1]};
        const float q[4] = {sp.rots[4 * (size_t)g], sp.rots[4 * (size_t)g + 1], sp.rots[4 * (size_t)g + 2], sp.rots[4 * (size_t)g + 3]};
        Derived d;
        derive_surfel(mu, sc, q, sp.opac[g], sp.mod, d);
        const float ax = sc[0] * d.f, ay = sc[1] * d.f;
        const float lx[4] = {-1.f, -1.f, 1.f, 1.f}, ly[4] = {1.f, -1.f, 1.f, -1.f};
        float v[12];
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int k = 0; k < 3; k++) v[3 * c + k] = (lx[c] * (d.tu[k] * ax) + ly[c] * (d.tv[k] * ay)) + mu[k];
        dst[4] = make_float4(v[0], v[1], v[2], v[3]); dst[5] = make_float4(v[4], v[5], v[6], v[7]);
        dst[6] = make_float4(v[8], v[9], v[10], v[11]); dst[7] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// fp64 Moeller-Trumbore on the fp32 corners, the oracle's stand-in for OptiX's ray / triangle test (oracle tri_hit(), operation
// for operation): the ray parameter of the hit, or -1
__device__ __forceinline__ double sp_tri_hit(const float* o, const float* d, const float* A, const float* B, const float* C)
{
    const double e1[3] = {(double)B[0] - A[0], (double)B[1] - A[1], (double)B[2] - A[2]};
    const double e2[3] = {(double)C[0] - A[0], (double)C[1] - A[1], (double)C[2] - A[2]};
    const double p[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
    const double det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (!(det != 0.0)) return -1.0;
    const double inv = 1.0 / det;
    const double s[3] = {(double)o[0] - A[0], (double)o[1] - A[1], (double)o[2] - A[2]};
    const double u = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) * inv;
    if (!(u >= 0.0 && u <= 1.0)) return -1.0;
    const double q[3] = {s[1] * e1[2] - s[2] * e1[1], s[2] * e1[0] - s[0] * e1[2], s[0] * e1[1] - s[1] * e1[0]};
    const double v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv;
    if (!(v >= 0.0 && u + v <= 1.0)) return -1.0;
    return (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
}

// One warp per ray.
__global__ void __launch_bounds__(128) k_sp_sort(BvhView bvh, FwdArgs a, WfBufs w, SpBufs sp)
{
    __shared__ unsigned long long s_keys[4][WF_HCAP];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned long long* keys = s_keys[wib];
    const SurfelRec* __restrict__ rec_g = bvh.rec_g;
    for (int s = blockIdx.x * 4 + wib; s < a.R; s += gridDim.x * 4) {
        const int r = w.order ? w.order[s] : s;
        const int n = sp.ccnt[r];
        if (n <= 0) {                                              // empty, handed to the fallback, or longer than WF_HCAP:
            const int hc = w.hit_count[r];                         // the latter are sorted by k_wf_sort_big and walked by k_sp_big
            if (!(hc & WF_TAINT) && hc > w.hcap) wf_claim_area(w, r, hc, lane);        // k_sp_slots reads ov_base: -1 = fallback
            else if (lane == 0 && !(hc & WF_TAINT) && hc > WF_HCAP) w.big_list[atomicAdd(w.counts + 10, 1)] = r;
            continue;
        }
        const long long base = sp.cbase[r];
        if ((base + n) * (sp.tri ? 2 : 1) > sp.capacity) {                              // the stream is full: this ray takes the per-ray fallback
            if (lane == 0) atomicOr(w.hit_count + r, WF_TAINT);
            continue;
        }
        const unsigned long long* bin = w.bins + (size_t)r * w.hcap;
        float4* dst = sp.srec + (sp.tri ? 8 : 4) * (size_t)base;
        // Up to 64 candidates (nearly every ray): the network runs on 32-bit keys — the depth's bits with the low 5 (6) replaced by
        // the candidate's slot — at half the shuffles, compares and selects of 64-bit (t, id) keys; the full key is then fetched
        // from the lane that holds it. The stream comes out sorted up to 2^-17 t, which pass A's margin allows for (wf_margin()):
        // its k-buffer orders the slots of a round by the exact (t', id) anyway.
        if (n <= 32) {
            const unsigned long long k = lane < n ? bin[lane] : LRT_KEY_EMPTY;
            unsigned k32 = lane < n ? (((unsigned)(k >> 32) & ~31u) | (unsigned)lane) : 0xffffffffu;
            k32 = warp_sort32(k32, lane);
            const unsigned long long ks = __shfl_sync(FULL, k, (int)(k32 & 31u));
            if (lane < n) sp_emit(rec_g, sp, dst, lane, ks);
            continue;
        }
        if (n <= 64) {
            const unsigned long long k0 = bin[lane], k1 = lane + 32 < n ? bin[lane + 32] : LRT_KEY_EMPTY;
            unsigned a0 = ((unsigned)(k0 >> 32) & ~63u) | (unsigned)lane;
            unsigned a1 = lane + 32 < n ? (((unsigned)(k1 >> 32) & ~63u) | (unsigned)(lane + 32)) : 0xffffffffu;
            warp_sort64(a0, a1, lane);
            {
                const int s0 = (int)(a0 & 63u), s1 = (int)(a1 & 63u);
                const unsigned long long x0 = __shfl_sync(FULL, k0, s0 & 31), y0 = __shfl_sync(FULL, k1, s0 & 31);
                const unsigned long long x1 = __shfl_sync(FULL, k0, s1 & 31), y1 = __shfl_sync(FULL, k1, s1 & 31);
                sp_emit(rec_g, sp, dst, lane, (s0 >> 5) ? y0 : x0);
                if (lane + 32 < n) sp_emit(rec_g, sp, dst, lane + 32, (s1 >> 5) ? y1 : x1);
            }
            continue;
        }
        if (n <= 128) {
            unsigned long long k[4];
#pragma unroll
            for (int j = 0; j < 4; j++) k[j] = lane + 32 * j < n ? bin[lane + 32 * j] : LRT_KEY_EMPTY;
            warp_sort_regs<4>(k, lane);
#pragma unroll
            for (int j = 0; j < 4; j++) if (lane + 32 * j < n) sp_emit(rec_g, sp, dst, lane + 32 * j, k[j]);
            continue;
        }
        if (n <= 256) {
            unsigned long long k[8];
#pragma unroll
            for (int j = 0; j < 8; j++) k[j] = lane + 32 * j < n ? bin[lane + 32 * j] : LRT_KEY_EMPTY;
            warp_sort_regs<8>(k, lane);
#pragma unroll
            for (int j = 0; j < 8; j++) if (lane + 32 * j < n) sp_emit(rec_g, sp, dst, lane + 32 * j, k[j]);
            continue;
        }
        const int m = WF_HCAP;                                     // 257..512 candidates: this warp's slice of shared memory
        for (int i = lane; i < m; i += 32) keys[i] = i < n ? bin[i] : LRT_KEY_EMPTY;
        __syncwarp(FULL);
        for (int size = 2; size <= m; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = lane; i < (m >> 1); i += 32) {
                    const int lo = ((i / stride) * stride * 2) + (i % stride), hi = lo + stride;
                    const unsigned long long x = keys[lo], y = keys[hi];
                    const bool up = ((lo & size) == 0);
                    if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
                }
                __syncwarp(FULL);
            }
        }
        for (int i = lane; i < n; i += 32) sp_emit(rec_g, sp, dst, i, keys[i]);
        __syncwarp(FULL);
    }
}

// ---- fused form (default, LRT_OPT_SPLIT_FUSED = 1): k_sp_sort and pass A as ONE kernel with one WARP per ray -------------------
// The two-kernel form moves every candidate's 64-byte record three times (gather, stream write, stream read: 0.87 GB per frame
// that SURVEY 8d does not know) so that a lone thread can walk them without dependent gathers — and that thread still waits on
// memory most of the time (17 % of the warps resident, 27 % issue slots used: profiles/r2_r_top_kernels_ncu.txt). Here the warp
// that sorted a ray's bin keeps the sorted keys in shared memory and runs the reference's rounds itself: per step the 32 lanes
// gather and test 32 candidates (exact test from the re-based origin AND the blending opacity, arithmetic of k_sp_slots), the
// accepted ones are compacted in bin order into the round's slot array, and the warp folds the 16 slots in order. No record is
// written anywhere; 32 gathers are in flight per warp instead of 4 per thread. Near-tie reordering (re-based depths out of bin
// order) falls back to a warp sort of the collected slots. Outputs are those of k_sp_slots, bit for bit.
__device__ __forceinline__ void sp_cp_async8(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void sp_cp_async4(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void sp_cp_async_wait_all();

__device__ __forceinline__ void sp_test_candidate(const SurfelRec* __restrict__ rec, unsigned long long ck, const RaySetup& rs, const FwdRay& q,
                                                  unsigned long long& key, float& alpha)
{
    key = LRT_KEY_EMPTY; alpha = 0.0f;
    const int g = (int)(unsigned)(ck & 0xffffffffull);
    const float4 a0 = ld_f4(&rec[g].r0), a1 = ld_f4(&rec[g].r1), a2 = ld_f4(&rec[g].r2), a3 = ld_f4(&rec[g].r3);
    // quad_hit(), operation for operation
    const float c0 = a0.x - rs.ox, c1 = a0.y - rs.oy, c2 = a0.z - rs.oz;
    const float den = a3.x * rs.dx + a3.y * rs.dy + a3.z * rs.dz;
    const float num = a3.x * c0 + a3.y * c1 + a3.z * c2;
    const float t = num / den;
    if (!(t > 0.0f)) return;
    {
        const float r0 = (rs.ox + t * rs.dx) - a0.x, r1 = (rs.oy + t * rs.dy) - a0.y, r2 = (rs.oz + t * rs.dz) - a0.z;
        const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;
        const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
        if (!(fabsf(u) <= a0.w && fabsf(v) <= a0.w) || !(t < LRT_TMAX)) return;
    }
    key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)g;
    // its opacity, as fwd_shade_round() computes it (forward.cu:212-251)
    const float dpt = t + q.base;
    const float x0 = q.o[0] + dpt * q.d[0], x1 = q.o[1] + dpt * q.d[1], x2 = q.o[2] + dpt * q.d[2];
    const float r0 = x0 - a0.x, r1 = x1 - a0.y, r2 = x2 - a0.z;
    const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;
    const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
    const float cosv = -((a0.x - q.o[0]) * a3.x + (a0.y - q.o[1]) * a3.y + (a0.z - q.o[2]) * a3.z);
    const float rho = u * u + v * v;
    const float power = -0.5f * rho;
    alpha = (cosv == 0.0f || power > 0.0f) ? 0.0f : fminf(LRT_ALPHA_MAX, a1.w * expf(power));
}

// warp-wide bitonic sort of one (64-bit key, float value) pair per lane, ascending by key
__device__ __forceinline__ void warp_sort32_kv(unsigned long long& k, float& v, int lane)
{
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k, stride);
            const float ov = __shfl_xor_sync(0xffffffffu, v, stride);
            const bool take_min = (((lane & size) == 0) == ((lane & stride) == 0));
            const bool other = take_min ? (ok < k) : (k < ok);
            k = other ? ok : k; v = other ? ov : v;
        }
    }
}

__device__ __forceinline__ void sp_slots_ray(const unsigned long long* keys, int n, int nw, int r, float em, const BvhView& bvh, const FwdArgs& a,
                                             const WfBufs& w, unsigned long long* sk, float* sa, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const SurfelRec* __restrict__ rec = bvh.rec_g;
    FwdRay q;
    fwd_ray_init(q, r, a);
    for (;;) {
        RaySetup rs;
        ray_setup(rs, q.o, q.d, q.base);
        const float thr = q.base - 2.0f * wf_margin(q.base, em);
        int pos = nw;                                              // first candidate of the sorted part at or beyond thr (warp-uniform lower bound)
        if (q.base > 0.0f) {                                       // first round: thr < 0 and every sorted key is > 0
            int lo = nw, hi = n;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__uint_as_float((unsigned)(keys[mid] >> 32)) < thr) lo = mid + 1; else hi = mid;
            }
            pos = lo;
        }
        // the round's slots: accepted candidates in bin order, compacted into sk / sa. Wild candidates (keys 0 .. nw) first.
        int have = 0;                                              // accepted so far (may exceed 32: then only counted)
        int have_wild = 0;                                         // ... of them from the wild segment (arbitrary depths, not in bin order)
        for (int seg = nw > 0 ? 0 : 1; seg < 2; seg++) {
            int i0 = seg ? pos : 0;
            const int i_hi = seg ? n : nw;
            if (seg) have_wild = have;
            for (; i0 < i_hi; i0 += 32) {
                if (seg && have - have_wild >= LRT_KBUF && have_wild + LRT_KBUF <= 32) {      // 16 accepted from the sorted part: does everything from here on lie safely behind the 16th of them?
                    const float t16 = __uint_as_float((unsigned)(sk[have_wild + LRT_KBUF - 1] >> 32)) + q.base;
                    if (__uint_as_float((unsigned)(keys[i0] >> 32)) - t16 > wf_margin(t16, em)) break;
                }
                unsigned long long key = LRT_KEY_EMPTY; float alpha = 0.0f;
                if (i0 + lane < i_hi) sp_test_candidate(rec, keys[i0 + lane], rs, q, key, alpha);
                const unsigned vm = __ballot_sync(FULL, key != LRT_KEY_EMPTY);
                const int at = have + __popc(vm & ((1u << lane) - 1u));
                if (key != LRT_KEY_EMPTY && at < 32) { sk[at] = key; sa[at] = alpha; }
                have += __popc(vm);
                __syncwarp(FULL);
                if (have > 32) break;                              // more than 32 accepted: handled below
            }
            if (have > 32) break;
        }
        int nvalid = min(have, 32);
        // accepted keys must ascend (re-basing can swap near-ties, wild candidates arrive out of order): otherwise sort the
        // collected slots; more than 32 collected means a window of ties wider than the slot array: take the general path
        {
            const unsigned long long mine = lane < nvalid ? sk[lane] : LRT_KEY_EMPTY;
            const unsigned long long prev = __shfl_up_sync(FULL, mine, 1);
            const bool bad = lane > 0 && lane < nvalid && prev > mine;
            if (__ballot_sync(FULL, bad) != 0 || have > 32) {
                if (have > 32) {
                    // general path: every candidate from the round's start, 32 at a time, keeping the 32 smallest (t', id) with their
                    // opacities (bitonic merge); stops like the fast path, on the true 16th
                    unsigned long long best = LRT_KEY_EMPTY; float bal = 0.0f;
                    for (int seg = nw > 0 ? 0 : 1; seg < 2; seg++) {
                        const int i_hi = seg ? n : nw;
                        for (int i0 = seg ? pos : 0; i0 < i_hi; i0 += 32) {
                            if (seg) {
                                const unsigned long long k16 = __shfl_sync(FULL, best, LRT_KBUF - 1);
                                if (k16 != LRT_KEY_EMPTY) {
                                    const float t16 = __uint_as_float((unsigned)(k16 >> 32)) + q.base;
                                    if (__uint_as_float((unsigned)(keys[i0] >> 32)) - t16 > wf_margin(t16, em)) break;
                                }
                            }
                            unsigned long long key = LRT_KEY_EMPTY; float alpha = 0.0f;
                            if (i0 + lane < i_hi) sp_test_candidate(rec, keys[i0 + lane], rs, q, key, alpha);
                            warp_sort32_kv(key, alpha, lane);
                            const unsigned long long rk = __shfl_sync(FULL, key, 31 - lane);
                            const float ra = __shfl_sync(FULL, alpha, 31 - lane);
                            if (rk < best) { best = rk; bal = ra; }                      // bitonic: element-wise minimum of (best, reversed new)
                            warp_sort32_kv(best, bal, lane);
                        }
                    }
                    sk[lane] = best; sa[lane] = bal;
                    nvalid = __popc(__ballot_sync(FULL, best != LRT_KEY_EMPTY));
                } else {
                    unsigned long long k = mine; float v = lane < nvalid ? sa[lane] : 0.0f;
                    warp_sort32_kv(k, v, lane);
                    sk[lane] = k; sa[lane] = v;
                }
                __syncwarp(FULL);
            }
        }
        // the round's compositing (fwd_shade_round()) without the colour. Lane i owns slot i: which slots contribute (:214, :220-224,
        // alpha >= 1/255) is decided by all lanes at once; only the transmittance / depth chain over the contributing slots is
        // walked in order (one broadcast load and six float operations per hit instead of the whole per-slot body on every lane:
        // that loop was 61 % of the kernel's instructions, profiles/r2_u_warp_ncu.txt); every owner then commits its own hit.
        const int nr = nvalid < LRT_KBUF ? nvalid : LRT_KBUF;
        bool terminated = false;
        {
            const unsigned long long key = lane < nr ? sk[lane] : 0ull;
            const float alpha = lane < nr ? sa[lane] : 0.0f;
            const int g = (int)(unsigned)(key & 0xffffffffull);
            const float dpt = __uint_as_float((unsigned)(key >> 32)) + q.base;           // forward.cu:212
            const bool okd = lane < nr && !(dpt < LRT_MIN_T);                            // :214
            const unsigned okm = __ballot_sync(FULL, okd);
            // :220-224: `last` is the id of the nearest slot in front that passed the depth test (a skipped duplicate leaves it unchanged)
            const unsigned front = okm & ((1u << lane) - 1u);
            const int gfront = __shfl_sync(FULL, g, front ? 31 - __clz(front) : 0);
            const bool contrib = okd && g != (front ? gfront : q.last) && !(alpha < 1.0f / 255.0f);
            const unsigned cm = __ballot_sync(FULL, contrib);
            // the contributing slots' (alpha, depth), compacted in order into this warp's slot array (every lane holds its own slot
            // in registers by now); the chain below reads them as broadcasts and leaves the transmittance in front of each hit
            const int mine = __popc(cm & ((1u << lane) - 1u));              // this lane's slot among the contributing ones
            const int nc = __popc(cm);
            float2* sc = reinterpret_cast<float2*>(sk);
            __syncwarp(FULL);
            if (contrib) sc[mine] = make_float2(alpha, dpt);
            __syncwarp(FULL);
            int stop = nr;                                                   // slot whose testT fell below T_MIN (:253-257)
            int ndone = nc;                                                  // contributing slots in front of it
            for (int i = 0; i < nc; i++) {
                const float2 v = sc[i];
                q.testT = q.T * (1.0f - v.x);
                if (q.testT < LRT_T_MIN) { terminated = true; ndone = i; break; }
                const float wgt = v.x * q.T;
                q.Dp += wgt * v.y; q.W += wgt;
                if (lane == 0) sa[i] = q.T;
                q.T = q.testT;
            }
            __syncwarp(FULL);
            if (terminated) {                                                // the original slot index of the terminating hit
                unsigned m = cm;
                for (int i = 0; i < ndone; i++) m &= m - 1;
                stop = __ffs(m) - 1;
            }
            const float myT = (contrib && mine < ndone) ? sa[mine] : 0.0f;
            q.nslots += terminated ? stop + 1 : nr;
            const unsigned done = cm & ((1u << stop) - 1u);                              // stop <= 16: the contributing slots in front of it
            if (contrib && lane < stop) {
                atomicAdd(a.accum_w + g, alpha * myT);                                   // :272
                const int at_k = q.ncontrib + __popc(done & ((1u << lane) - 1u));
                if (at_k < a.cap) {
                    const size_t at = (size_t)at_k * a.R + q.r;
                    a.hit_gidx[at] = g;
                    a.hit_t[at] = dpt;
                    reinterpret_cast<float*>(a.hit_aux + at)[0] = alpha;                  // colour: k_sp_colour
                }
            }
            q.ncontrib += __popc(done);
            if (!terminated && nr > 0) {
                q.dpt = __shfl_sync(FULL, dpt, nr - 1);
                if (okm) q.last = __shfl_sync(FULL, g, 31 - __clz(okm));
            }
        }
        __syncwarp(FULL);
        if (terminated || q.testT < LRT_T_MIN || nvalid < LRT_KBUF) break;                // :282-285
        q.base = (float)((double)q.dpt + LRT_STEP_EPS);                                   // :288
    }
    if (lane == 0) {
        float* op = a.out + (size_t)LRT_NCH * q.r;                                        // :296-305 minus the colour channels (k_sp_fold)
        op[3] = q.Dp; op[4] = q.W; op[5] = 0.f; op[6] = 0.f; op[7] = 0.f; op[8] = q.T;
        a.hit_cnt[q.r] = q.ncontrib;
        if (a.slot_cnt) a.slot_cnt[q.r] = q.nslots;
        if (q.ncontrib > a.cap) w.ov_list[atomicAdd(w.counts + 12, 1)] = q.r;             // list truncated: the ray is redone per ray
    }
}

#ifndef LRT_WARP_MIN_BLOCKS
#define LRT_WARP_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(128, LRT_WARP_MIN_BLOCKS) k_sp_warp(BvhView bvh, FwdArgs a, WfBufs w)
{
    __shared__ unsigned long long s_keys[4][WF_HCAP];
    __shared__ unsigned long long s_sk[4][32];
    __shared__ float s_sa[4][32];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned long long* keys = s_keys[wib];
    // the next ray's candidate count and the first 32 keys of its bin travel to shared memory (cp.async: no destination register,
    // nothing for the compiler to spill) while this ray is processed: the two dependent global loads in front of every sort were
    // 4.6 % of the kernel's stall samples (profiles/r2_z_warp_ncu.txt)
    __shared__ unsigned long long s_pf[4][32];
    __shared__ int s_hc[4];
    int r = blockIdx.x * 4 + wib;
    if (r < a.R) {
        sp_cp_async8(&s_pf[wib][lane], w.bins + (size_t)r * w.hcap + lane);
        if (lane == 0) sp_cp_async4(&s_hc[wib], w.hit_count + r);
    }
    for (; r < a.R; r += gridDim.x * 4) {
        sp_cp_async_wait_all();
        __syncwarp(FULL);
        const int hc = s_hc[wib];
        const unsigned long long k_first = s_pf[wib][lane];
        __syncwarp(FULL);
        {
            const int r2 = r + gridDim.x * 4;
            if (r2 < a.R) {
                sp_cp_async8(&s_pf[wib][lane], w.bins + (size_t)r2 * w.hcap + lane);
                if (lane == 0) sp_cp_async4(&s_hc[wib], w.hit_count + r2);
            }
        }
        if ((hc & WF_TAINT) || (hc > w.hcap && !wf_claim_area(w, r, hc, lane))) {      // dropped candidates / no room for its overflow: per-ray fallback
            if (lane == 0) { w.fb_list[atomicAdd(w.counts + 8, 1)] = r; a.hit_cnt[r] = 0; }
            continue;
        }
        if (hc > w.hcap) continue;                                 // listed by wf_claim_area: sorted by k_wf_sort_big, walked by k_sp_big
        if (hc > WF_HCAP) {                                        // sorted by k_wf_sort_big, walked by k_sp_big
            if (lane == 0) w.big_list[atomicAdd(w.counts + 10, 1)] = r;
            continue;
        }
        const int n = hc;
        const unsigned long long* bin = w.bins + (size_t)r * w.hcap;
        // ---- sort the bin into this warp's slice of shared memory (register networks as in k_sp_sort)
        if (n <= 32) {
            const unsigned long long k = lane < n ? k_first : LRT_KEY_EMPTY;
            unsigned k32 = lane < n ? (((unsigned)(k >> 32) & ~31u) | (unsigned)lane) : 0xffffffffu;
            k32 = warp_sort32(k32, lane);
            keys[lane] = __shfl_sync(FULL, k, (int)(k32 & 31u));
        } else if (n <= 64) {
            const unsigned long long k0 = k_first, k1 = lane + 32 < n ? bin[lane + 32] : LRT_KEY_EMPTY;
            unsigned a0 = ((unsigned)(k0 >> 32) & ~63u) | (unsigned)lane;
            unsigned a1 = lane + 32 < n ? (((unsigned)(k1 >> 32) & ~63u) | (unsigned)(lane + 32)) : 0xffffffffu;
            warp_sort64(a0, a1, lane);
            const int s0 = (int)(a0 & 63u), s1 = (int)(a1 & 63u);
            const unsigned long long x0 = __shfl_sync(FULL, k0, s0 & 31), y0 = __shfl_sync(FULL, k1, s0 & 31);
            const unsigned long long x1 = __shfl_sync(FULL, k0, s1 & 31), y1 = __shfl_sync(FULL, k1, s1 & 31);
            keys[lane] = (s0 >> 5) ? y0 : x0;
            keys[lane + 32] = a1 == 0xffffffffu ? LRT_KEY_EMPTY : ((s1 >> 5) ? y1 : x1);
        } else if (n <= 128) {
            unsigned long long k[4];
#pragma unroll
            for (int j = 0; j < 4; j++) k[j] = lane + 32 * j < n ? bin[lane + 32 * j] : LRT_KEY_EMPTY;
            warp_sort_regs<4>(k, lane);
#pragma unroll
            for (int j = 0; j < 4; j++) keys[lane + 32 * j] = k[j];
        } else if (n <= 256) {
            unsigned long long k[8];
#pragma unroll
            for (int j = 0; j < 8; j++) k[j] = lane + 32 * j < n ? bin[lane + 32 * j] : LRT_KEY_EMPTY;
            warp_sort_regs<8>(k, lane);
#pragma unroll
            for (int j = 0; j < 8; j++) keys[lane + 32 * j] = k[j];
        } else {
            const int m = WF_HCAP;
            for (int i = lane; i < m; i += 32) keys[i] = i < n ? bin[i] : LRT_KEY_EMPTY;
            __syncwarp(FULL);
            for (int size = 2; size <= m; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int i = lane; i < (m >> 1); i += 32) {
                        const int lo = ((i / stride) * stride * 2) + (i % stride), hi = lo + stride;
                        const unsigned long long x = keys[lo], y = keys[hi];
                        const bool up = ((lo & size) == 0);
                        if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
                    }
                    __syncwarp(FULL);
                }
            }
        }
        __syncwarp(FULL);
        sp_slots_ray(keys, n, min(w.nwild[r], n), r, __int_as_float(w.emax[r]), bvh, a, w, s_sk[wib], s_sa[wib], lane);
        __syncwarp(FULL);
    }
}

// The rays with more than WF_HCAP candidates (a ray skimming a wall or the side of a vehicle: thousands), whose bins
// k_wf_sort_big has sorted in place: one WARP per ray walks the sorted keys in global memory and does the whole job — slots,
// colour, fold, hit lists (wf_shade_ray). Pass A skips these rays; passes B and C find complete hit records and reproduce the
// same colours and sums. Replaces what used to be the "fallback cliff": such a ray cost a lone thread milliseconds.
__global__ void __launch_bounds__(128) k_sp_big(BvhView bvh, FwdArgs a, WfBufs w)
{
    const int lane = threadIdx.x & 31;
    const int nbig = min(w.counts[10], a.R);
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nbig; b += (gridDim.x * blockDim.x) >> 5) {
        const int r = w.big_list[b];
        const int n = w.hit_count[r] & (WF_TAINT - 1);             // beyond hcap: the ray's area of the overflow arena
        const float em = w.nwild[r] > 0 ? __int_as_float(0x7f800000) : __int_as_float(w.emax[r]);
        wf_shade_ray(wf_big_keys(w, r, n), n, r, em, bvh, a, lane);
        __syncwarp(0xffffffffu);
    }
}

#ifndef LRT_SLOTS_MIN_BLOCKS
#define LRT_SLOTS_MIN_BLOCKS 4
#endif
#define SP_BATCH 4                                                  // candidates staged per wait
#define SP_SLOTS_SMEM (128 * (LRT_KBUF * (8 + 4) + SP_BATCH * 64)) // k-buffer keys + opacities + staging, per block of 128 threads

__device__ __forceinline__ void sp_cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void sp_cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// Pass A. The walk of k_wf_composite2 over the ray's candidates, now a contiguous stream of records sorted by (t from o, id);
// everything that needs an SH row is gone. A thread stages SP_BATCH candidates (all four 16-byte parts of each) into its own
// column of shared memory with cp.async and waits ONCE per batch: ncu showed the first form of this kernel (records fetched into
// registers, the inner parts on demand) waiting five times per four candidates, every wait a full L2 / DRAM round trip
// (profiles/r2_b_split_first_ncu.txt). Columns are [slot][thread]: consecutive lanes touch consecutive 16-byte words (no bank
// conflicts) and no thread ever reads another thread's column, so no block-level barrier is needed.
// TRI (LRT_OPT_TRIANGLE_DEPTH): a candidate's hits come from the reference's literal proxy — the two triangles (v0,v1,v2), (v2,v3,v1)
// over the fp32-rounded corners, fp64 Moeller-Trumbore like the oracle's ORC_TRIANGLES mode — instead of the analytic quad: keys are
// (t', primitive id = 2 g + triangle), a candidate can give two slots (a ray through the shared diagonal), the depth of a hit is
// the triangle's. Records are 128 bytes, staged two at a time.
template <bool TRI>
__global__ void __launch_bounds__(128, LRT_SLOTS_MIN_BLOCKS) k_sp_slots(FwdArgs a, WfBufs w, SpBufs sp)
{
    constexpr int RS = TRI ? 8 : 4;                                 // float4 per record
    constexpr int NB = TRI ? SP_BATCH / 2 : SP_BATCH;               // records staged per wait (same bytes either way)
    constexpr int NH = TRI ? 2 : 1;                                 // possible hits per candidate
    extern __shared__ __align__(16) unsigned char sp_smem[];
    float4 (*s_st)[128] = reinterpret_cast<float4 (*)[128]>(sp_smem);                                      // [SP_BATCH * 4][128] staging
    unsigned long long (*s_kb)[128] = reinterpret_cast<unsigned long long (*)[128]>(sp_smem + 128 * SP_BATCH * 64);   // the round's slots, ascending (t', id)
    float (*s_al)[128] = reinterpret_cast<float (*)[128]>(sp_smem + 128 * (SP_BATCH * 64 + LRT_KBUF * 8));             // their blending opacities
    const int tx = threadIdx.x;
    for (int s = blockIdx.x * blockDim.x + tx; s < a.R; s += gridDim.x * blockDim.x) {
        const int r = w.order ? w.order[s] : s;
        const int hc = w.hit_count[r];
        if ((hc & WF_TAINT) || (hc > w.hcap && !(w.ov_pairs && w.ov_base[r] >= 0))) { w.fb_list[atomicAdd(w.counts + 8, 1)] = r; a.hit_cnt[r] = 0; continue; }
        if (hc > WF_HCAP) continue;                                // done by k_sp_big, one warp per ray
        const int n = hc;
        const float em = __int_as_float(w.emax[r]);
        const int nw = min(w.nwild[r], n);                         // wild candidates (key t = 0): the first nw of the sorted stream, tested in every round
        const float4* __restrict__ rec = sp.srec + RS * (size_t)sp.cbase[r];             // candidate i = rec[RS i .. RS i + RS - 1]
        FwdRay q;
        fwd_ray_init(q, r, a);
        int pos = nw;                                              // first candidate of the sorted part that can still matter
        int i_end = nw;                                            // where the previous round's scan stopped: everything from there on lies beyond thr
        for (;;) {
            RaySetup rs;
            ray_setup(rs, q.o, q.d, q.base);
            const float thr = q.base - 2.0f * wf_margin(q.base, em);
            {   // first candidate at or beyond thr (the stream is sorted by t). It lies a few entries in front of where the
                // previous round stopped: walk back from there eight independent loads at a time (one wait per eight).
                int hi = i_end;
                while (hi > pos) {
                    float tv[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) tv[k] = hi - 1 - k >= pos ? __ldg(reinterpret_cast<const float*>(rec + RS * (hi - 1 - k) + 3) + 3) : -1.0f;
                    int back = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) if (back == k && hi - 1 - k >= pos && !(tv[k] < thr)) back = k + 1;
                    hi -= back;
                    if (back < 8) break;
                }
                pos = hi;
            }
            i_end = n;
            int cnt = 0;
            unsigned long long klast = 0ull;                       // s_kb[cnt - 1]
            bool done = false;
            for (int seg = nw > 0 ? 0 : 1; seg < 2; seg++) {       // segment 0: the wild candidates, no window; segment 1: the sorted rest
            const int i_lo = seg ? pos : 0, i_hi = seg ? n : nw;
            for (int i0 = i_lo; i0 < i_hi && !done; i0 += NB) {
                {
                    const float4* src = rec + RS * (size_t)i0;
                    const int nv = RS * min(NB, i_hi - i0);
#pragma unroll
                    for (int v = 0; v < RS * NB; v++) if (v < nv) sp_cp_async16(&s_st[v][tx], src + v);
                    sp_cp_async_wait_all();
                }
                // phase 1, branch-free: the exact test and the opacity of all SP_BATCH staged candidates, independent of each other
                // (the instruction streams interleave: this kernel runs few warps per scheduler and every dependent chain of
                // shared-memory loads, divisions and expf would otherwise be paid in full, one candidate after the other)
                float t0v[NB], alv[NB * NH];
                unsigned long long keyv[NB * NH];
                bool hitv[NB * NH];
#pragma unroll
                for (int k = 0; k < NB; k++) {
                    const float4 a0 = s_st[RS * k][tx], a1 = s_st[RS * k + 1][tx], a2 = s_st[RS * k + 2][tx], a3 = s_st[RS * k + 3][tx];
                    t0v[k] = a3.w;
                    float th[NH];
                    if (TRI) {
                        const float4 p4 = s_st[RS * k + 4][tx], p5 = s_st[RS * k + 5][tx], p6 = s_st[RS * k + 6][tx];
                        const float v0[3] = {p4.x, p4.y, p4.z}, v1[3] = {p4.w, p5.x, p5.y}, v2[3] = {p5.z, p5.w, p6.x}, v3[3] = {p6.y, p6.z, p6.w};
                        const float oo[3] = {rs.ox, rs.oy, rs.oz}, dd[3] = {rs.dx, rs.dy, rs.dz};
                        // prim 2g = (v0,v1,v2), prim 2g+1 = (v2,v3,v1): primitive_utils.py:212-221; oracle test_gauss()
                        const double ta = sp_tri_hit(oo, dd, v0, v1, v2), tb = sp_tri_hit(oo, dd, v2, v3, v1);
                        th[0] = (float)ta; th[NH - 1] = (float)tb;
                        hitv[NH * k] = ta > 0.0 && th[0] > 0.0f && th[0] < LRT_TMAX;
                        hitv[NH * k + NH - 1] = tb > 0.0 && th[NH - 1] > 0.0f && th[NH - 1] < LRT_TMAX;
                    } else {
                        // quad_hit(), operation for operation
                        const float c0 = a0.x - rs.ox, c1 = a0.y - rs.oy, c2 = a0.z - rs.oz;
                        const float den = a3.x * rs.dx + a3.y * rs.dy + a3.z * rs.dz;
                        const float num = a3.x * c0 + a3.y * c1 + a3.z * c2;
                        const float t = num / den;
                        bool h = t > 0.0f;
                        const float r0 = (rs.ox + t * rs.dx) - a0.x, r1 = (rs.oy + t * rs.dy) - a0.y, r2 = (rs.oz + t * rs.dz) - a0.z;
                        const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;
                        const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
                        h = h && (fabsf(u) <= a0.w && fabsf(v) <= a0.w) && (t < LRT_TMAX);
                        hitv[k] = h; th[0] = t;
                    }
#pragma unroll
                    for (int hh = 0; hh < NH; hh++) {
                        const float t = th[hh];
                        keyv[NH * k + hh] = ((unsigned long long)__float_as_uint(t) << 32) |
                                            (TRI ? 2u * (unsigned)__float_as_int(a2.w) + (unsigned)hh : (unsigned)__float_as_int(a2.w));
                        // its opacity, as fwd_shade_round() computes it (forward.cu:212-251)
                        const float dpt = t + q.base;
                        const float x0 = q.o[0] + dpt * q.d[0], x1 = q.o[1] + dpt * q.d[1], x2 = q.o[2] + dpt * q.d[2];
                        const float r0 = x0 - a0.x, r1 = x1 - a0.y, r2 = x2 - a0.z;
                        const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;
                        const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
                        const float cosv = -((a0.x - q.o[0]) * a3.x + (a0.y - q.o[1]) * a3.y + (a0.z - q.o[2]) * a3.z);
                        const float rho = u * u + v * v;
                        const float power = -0.5f * rho;
                        alv[NH * k + hh] = (cosv == 0.0f || power > 0.0f) ? 0.0f : fminf(LRT_ALPHA_MAX, a1.w * expf(power));
                    }
                }
                // phase 2, in order: window stop, k-buffer
#pragma unroll
                for (int kk = 0; kk < NB * NH; kk++) {
                    const int k = kk / NH;
                    if (i0 + k >= i_hi || done) continue;
                    if (seg && cnt == LRT_KBUF && kk % NH == 0) {
                        const float t16 = __uint_as_float((unsigned)(klast >> 32)) + q.base;
                        if (t0v[k] - t16 > wf_margin(t16, em)) { done = true; i_end = i0 + k; continue; }
                    }
                    if (!hitv[kk]) continue;
                    const unsigned long long key = keyv[kk];
                    const float alpha = alv[kk];
                    if (cnt == LRT_KBUF && key >= klast) continue;                              // behind the current 16th
                    if (cnt < LRT_KBUF && (cnt == 0 || key > klast)) {                          // the usual case: append
                        s_kb[cnt][tx] = key; s_al[cnt][tx] = alpha; cnt++; klast = key;
                        continue;
                    }
                    if (cnt == LRT_KBUF) cnt--;                                                 // replaces the current 16th
                    int j = cnt;
                    while (j > 0 && s_kb[j - 1][tx] > key) { s_kb[j][tx] = s_kb[j - 1][tx]; s_al[j][tx] = s_al[j - 1][tx]; j--; }
                    s_kb[j][tx] = key; s_al[j][tx] = alpha;
                    cnt++;
                    klast = s_kb[cnt - 1][tx];
                }
            }
            }
            // the round's compositing (fwd_shade_round()) without the colour: weights, depth, transmittance, hit records
            bool terminated = false;
            for (int i4 = 0; i4 < cnt && !terminated; i4 += 4) {
                unsigned long long kq[4]; float aq[4];
#pragma unroll
                for (int j = 0; j < 4; j++) { kq[j] = s_kb[(i4 + j) & (LRT_KBUF - 1)][tx]; aq[j] = s_al[(i4 + j) & (LRT_KBUF - 1)][tx]; }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (i4 + j >= cnt || terminated) continue;
                    const unsigned long long key = kq[j];
                    const int g = TRI ? (int)((unsigned)(key & 0xffffffffull) >> 1) : (int)(unsigned)(key & 0xffffffffull);     // gidx = pidx / 2
                    q.nslots++;
                    q.dpt = __uint_as_float((unsigned)(key >> 32)) + q.base;              // forward.cu:212
                    if (q.dpt < LRT_MIN_T) continue;                                      // :214
                    if (g == q.last) continue;                                            // :220-224
                    q.last = g;
                    const float alpha = aq[j];
                    if (alpha < 1.0f / 255.0f) continue;
                    q.testT = q.T * (1.0f - alpha);
                    if (q.testT < LRT_T_MIN) { terminated = true; continue; }             // :253-257
                    const float wgt = alpha * q.T;
                    q.Dp += wgt * q.dpt; q.W += wgt;
                    atomicAdd(a.accum_w + g, wgt);                                        // :272
                    if (q.ncontrib < a.cap) {
                        const size_t at = (size_t)q.ncontrib * a.R + q.r;
                        a.hit_gidx[at] = g;
                        a.hit_t[at] = q.dpt;
                        reinterpret_cast<float*>(a.hit_aux + at)[0] = alpha;              // colour: k_sp_colour
                    }
                    q.ncontrib++;
                    q.T = q.testT;
                }
            }
            if (terminated || q.testT < LRT_T_MIN || cnt < LRT_KBUF) break;               // :282-285
            q.base = (float)((double)q.dpt + LRT_STEP_EPS);                               // :288
        }
        float* op = a.out + (size_t)LRT_NCH * q.r;                                        // :296-305 minus the colour channels (k_sp_fold)
        op[3] = q.Dp; op[4] = q.W; op[5] = 0.f; op[6] = 0.f; op[7] = 0.f; op[8] = q.T;
        a.hit_cnt[q.r] = q.ncontrib;
        if (a.slot_cnt) a.slot_cnt[q.r] = q.nslots;
        if (q.ncontrib > a.cap) w.ov_list[atomicAdd(w.counts + 12, 1)] = q.r;             // list truncated: the ray is redone per ray
    }
}

// colour from a row whose dc part sits in `dc` and whose rest part starts A floats into the 16-byte aligned window `wv`
// (same sums, in the same order, as sh_colour_stream_b() over the concatenated row)
template <int A>
__device__ __forceinline__ void sp_colour_window(int nb, const float* b, const float* dc, const float* wv, float* c)
{
    const int nf = 3 * nb;
    float r[3] = {b[0] * dc[0], b[0] * dc[1], b[0] * dc[2]};
#pragma unroll
    for (int idx = 3; idx < 48; idx++) {
        if (idx < nf) { const int j = idx / 3, ch = idx % 3; r[ch] = r[ch] + b[j] * wv[A + idx - 3]; }
    }
    c[0] = r[0] + 0.5f; c[1] = r[1] + 0.5f; c[2] = r[2] + 0.5f;
    if (c[0] < 0.0f) c[0] = -0.0f;     // see sh_colour
}

// Pass B: SH colour of every recorded hit (forward.cu:67-111, :261-266). Thread (x, y) owns ray x and its hits y, y + gridDim.y, ...
// MODE 1: concatenated (P, M, 3) rows, 16-byte aligned: twelve 128-bit loads. MODE 2: rows read in place from features_dc /
// features_rest (lrt_set_sh_parts): a `rest` row is 3 (M - 1) floats at a 4-byte aligned address, so the thread loads the twelve
// 16-byte words that cover it and picks its coefficients at one of four compile-time offsets. MODE 0: anything else, scalar loads.
#ifndef LRT_COLOUR_MIN_BLOCKS
#define LRT_COLOUR_MIN_BLOCKS 4
#endif
template <int MODE>
__global__ void __launch_bounds__(256, MODE == 2 ? 3 : LRT_COLOUR_MIN_BLOCKS) k_sp_colour(FwdArgs a)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const int cnt = min(a.hit_cnt[r], a.cap);
    if ((int)blockIdx.y >= cnt) return;
    float d[3], dirn[3];
#pragma unroll
    for (int k = 0; k < 3; k++) d[k] = a.ray_d[3 * (size_t)r + k];
    const float dl = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);                     // fwd_ray_init()
#pragma unroll
    for (int k = 0; k < 3; k++) dirn[k] = d[k] / dl;
    float sb[16];
    const int nb = sh_basis(a.D, dirn, sb);
    int g_nx = a.hit_gidx[(size_t)blockIdx.y * a.R + r];
    for (int k = blockIdx.y; k < cnt; k += gridDim.y) {
        const size_t at = (size_t)k * a.R + r;
        const int g = g_nx;
        if (k + (int)gridDim.y < cnt) {                            // the thread's next hit: its SH row starts moving towards L2 now
            g_nx = a.hit_gidx[at + (size_t)gridDim.y * a.R];
            if (MODE == 1) {
                const char* row = reinterpret_cast<const char*>(a.shs + (size_t)g_nx * a.M * 3);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                if (nb > 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
            }
        }
        float c[3];
        if (MODE == 1) {
            sh_colour_stream_b(nb, sb, a.shs + (size_t)g * a.M * 3, c);
        } else if (MODE == 2) {
            int j;
            const ShPartDev& p = sh_find(a.sh_tab, g, j);
            {
                const int rf = 3 * (a.sh_tab->M - 1);
                const size_t f0 = (size_t)rf * j;
                const int al = (int)(f0 & 3);
                const float* w0 = p.rest + (f0 - al);
                const float4* p4 = reinterpret_cast<const float4*>(w0);
                const float* dcp = p.dc + 3 * (size_t)j;
                const float dc[3] = {ld_f(dcp), ld_f(dcp + 1), ld_f(dcp + 2)};
                const int n4 = (al + 3 * nb - 3 + 3) >> 2;
                const int left = j + 1 == p.P ? al + rf : 64;      // floats of the window that exist (the tensor's last row: no read past its end)
                float wv[48];
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < n4) {
                        if (4 * i + 4 <= left) v = ld_f4(p4 + i);
                        else {
                            if (4 * i < left) v.x = ld_f(w0 + 4 * i);
                            if (4 * i + 1 < left) v.y = ld_f(w0 + 4 * i + 1);
                            if (4 * i + 2 < left) v.z = ld_f(w0 + 4 * i + 2);
                        }
                    }
                    wv[4 * i] = v.x; wv[4 * i + 1] = v.y; wv[4 * i + 2] = v.z; wv[4 * i + 3] = v.w;
                }
                // bring the row to the front of the window: shift by 1 and by 2 floats as the bits of `al` say (predicated moves on
                // registers; a 4-way branch on `al` would run the whole evaluation once per alignment present in the warp)
#pragma unroll
                for (int i = 0; i < 47; i++) wv[i] = (al & 1) ? wv[i + 1] : wv[i];
#pragma unroll
                for (int i = 0; i < 46; i++) wv[i] = (al & 2) ? wv[i + 2] : wv[i];
                sp_colour_window<0>(nb, sb, dc, wv, c);
            }
        } else {
            float sh[48]; bool cl;
            load_sh_any(a, g, nb, sh);
            sh_colour<false>(a.D, dirn, sh, c, cl, nullptr);
        }
        float* ax = reinterpret_cast<float*>(a.hit_aux + at);
        ax[1] = c[0]; ax[2] = c[1]; ax[3] = c[2];
    }
}

// Pass C: the ordered fold of the colour channels (forward.cu:253-270, :296-298) over the recorded (alpha, colour).
__global__ void __launch_bounds__(128) k_sp_fold(FwdArgs a)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const int cnt = a.hit_cnt[r];
    if (cnt > a.cap) return;                                       // redone by k_wf_fallback_overflow
    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    for (int k0 = 0; k0 < cnt; k0 += 8) {                          // eight independent loads per wait
        float4 h[8];
#pragma unroll
        for (int j = 0; j < 8; j++) if (k0 + j < cnt) h[j] = a.hit_aux[(size_t)(k0 + j) * a.R + r];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (k0 + j < cnt) {
                const float testT = T * (1.0f - h[j].x);
                const float wgt = h[j].x * T;
                C0 += wgt * h[j].y; C1 += wgt * h[j].z; C2 += wgt * h[j].w;
                T = testT;
            }
        }
    }
    float* op = a.out + (size_t)LRT_NCH * r;
    op[0] = C0 + T * a.bg[0]; op[1] = C1 + T * a.bg[1]; op[2] = C2 + T * a.bg[2];
}
