// lrt_wavefront.cuh — forward "kernel D": breadth-first wavefront traversal + warp-per-ray compositing.
// Included by lrt_forward.cu inside its anonymous namespace (uses FwdArgs, FwdRay, G8Slot, g8_shade_slot).
//
// Why: per-ray traversal (kernels A/B/C) keeps a warp's lanes in different places of different trees:
// measured 9-15 active lanes per instruction and ~3-5e9 warp instructions per frame however the loop is
// organised. The reference semantics do not need a per-ray loop at all: a round's k-buffer is "the 16
// nearest proxy hits beyond the re-basing point", and the whole ray carries only ~40 proxy hits in this
// workload. So:
//   1. k_wf_level (one launch per hierarchy level, top-down): one thread per (ray, node) work item,
//      evaluates the node's 8 child boxes and appends one item per entered child with a warp-aggregated
//      atomic — every lane runs the same straight-line code, no state machine, no culling needed.
//   2. k_wf_leaf: one thread per (ray, leaf) item: 8 surfel boxes, exact quad tests, hits appended to the
//      ray's bin as 64-bit keys (t bits, Gaussian id).
//   3. k_wf_shade: one warp per ray: sorts the bin (shared memory bitonic), then replays the reference's
//      rounds exactly: every round re-tests a 32-candidate window of the sorted bin from the (re-based) origin o' = o + base d (same arithmetic as the per-ray kernels, so t',
//      the epsilon gap and duplicate suppression are reproduced bit for bit), shades 16 slots in parallel
//      and folds them in order with shuffles.
//   4. k_wf_fallback: rays the wavefront cannot guarantee (bin or work-list overflow, window guard) are
//      traced by the per-ray code path. Same results, just slower; normally a handful of rays.
#pragma once

#define WF_HCAP 512                 // bins up to this size are sorted in shared memory (and are all k_wf_shade can take)
#define WF_HCAP_MAX 8192            // bin capacity per ray in memory (a full, un-culled ray carries ~40 candidates on street scenes,
                                    // a few grazing rays several hundred, a ray skimming the side of a vehicle next to the sensor
                                    // thousands); beyond it the ray goes to the per-ray fallback, which costs MILLISECONDS per ray
                                    // (one thread re-walking the hierarchy once per 16 hits). Only touched lines cost traffic.
#define WF_TAINT 0x40000000         // hit_count flag: work item or hit dropped -> fallback
#define WF_WINDOW_MARGIN 1e-3f

struct WfBufs {
    RaySetup* rs;                   // (R) per-ray slab constants for base = 0
    uint2* list_a; uint2* list_b;   // ping-pong work lists {ray, node}
    int cap_items;
    int* counts;                    // [0..7] items per level, [8] fallback count, [9] heavy beam-grid items, [10] bins beyond 512 candidates,
                                    // [12] rays whose hit list overflowed in the split passes, [13] candidate records in the sorted stream
    int* hit_count;                 // (R) hits in bin | WF_TAINT
    int* emax;                      // (R) float bits: largest depth-error bound among the ray's candidates (0 below LRT_ERR_FLOOR)
    int* nwild;                     // (R) candidates whose bound hit the cap ("wild": depth numerically meaningless); their keys carry t = 0
    unsigned long long* bins;       // (R, hcap)
    int hcap;                       // bin capacity (<= WF_HCAP_MAX)
    int* fb_list;                   // (R) fallback ray ids
    int* big_list;                  // (R) rays whose bin holds more than WF_HCAP candidates: sorted by k_wf_sort_big, one block each
    int* ov_list;                   // (R) split passes: rays whose contributing-hit list overflowed `cap`
    int* ray_ids;                   // (R) identity, input of the by-length sort
    const int* order;               // (R) rays sorted by descending candidate count, or nullptr (tile order)
    // Candidates beyond a bin's capacity (split passes; nullptr elsewhere: such a ray then takes the per-ray fallback). They are
    // appended as (ray, key) pairs to one list; a ray that has some gets an AREA of the arena (k_sp_warp / k_sp_sort: its bin is
    // copied there, k_ov_scatter adds its pairs), which k_wf_sort_big sorts and k_sp_big walks like any long bin.
    uint4* ov_pairs; int ov_pair_cap;          // counts[14] = pairs appended
    unsigned long long* ov_area; int ov_area_cap;      // counts[15] = keys handed out (areas are padded to a power of two)
    int* ov_base;                   // (R) first key of the ray's area, -1: none (the ray takes the fallback)
    int* ov_fill;                   // (R) pairs of the ray already placed
};

// keys of a ray of the long-bin list: its bin, or its area of the overflow arena
__device__ __forceinline__ unsigned long long* wf_big_keys(const WfBufs& w, int r, int hc)
{
    return hc > w.hcap ? w.ov_area + w.ov_base[r] : w.bins + (size_t)r * w.hcap;
}

__global__ void __launch_bounds__(256) k_wf_setup(FwdArgs a, WfBufs w)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    float o[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { o[k] = a.ray_o[(size_t)r * a.ray_o_stride + k]; d[k] = a.ray_d[3 * (size_t)r + k]; }
    RaySetup rs;
    ray_setup(rs, o, d, 0.0f);
    w.rs[r] = rs;
    w.hit_count[r] = 0;
    w.emax[r] = 0; w.nwild[r] = 0;
    w.ray_ids[r] = r;
    if (w.ov_fill) w.ov_fill[r] = 0;
}

// One (ray, node) item per thread at `level` >= 1. in == nullptr: the implicit root list (every ray, node 0)
// in 4x8 tile order.
__global__ void __launch_bounds__(256) k_wf_level(BvhView bvh, FwdArgs a, WfBufs w, int level, const uint2* __restrict__ in,
                                                 const int* __restrict__ in_count, uint2* __restrict__ out, int* __restrict__ out_count)
{
    const int n_in = in ? min(*in_count, w.cap_items) : num_slots(a.R, a.grid_w);
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n_in; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        int ray = -1; unsigned node = 0;
        if (i < n_in) {
            if (in) { const uint2 it = in[i]; ray = (int)it.x; node = it.y; }
            else ray = slot_to_ray(i, a.R, a.grid_w);
        }
        unsigned m = 0;
        if (ray >= 0) {
            const RaySetup rs = w.rs[ray];
            int nearest;
            m = node_eval(bvh.nodes + bvh.level_off[level] + node, rs, LRT_TMAX, 0xffu, nearest);
        }
        // warp-aggregated append: exclusive scan of the per-lane child counts
        const int cnt = __popc(m);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        int wbase = 0;
        if (lane == 31) wbase = atomicAdd(out_count, total);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        int off = wbase + incl - cnt;
        if (cnt) {
            if (off + cnt <= w.cap_items) {
                while (m) { const int c = __ffs(m) - 1; m &= m - 1; out[off++] = make_uint2((unsigned)ray, node * 8u + c); }
            } else {
                atomicOr(w.hit_count + ray, WF_TAINT);          // work list full: this ray goes to the fallback
            }
        }
    }
}

// 8 slab tests against the byte-quantised boxes of a compact leaf.
__device__ __forceinline__ unsigned leafq_eval(const LeafQ* __restrict__ lq, const RaySetup& r)
{
    const uint4* p = reinterpret_cast<const uint4*>(lq);
    const uint4 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3), w4 = __ldg(p + 4);
    const float lox = __uint_as_float(w0.x), loy = __uint_as_float(w0.y), loz = __uint_as_float(w0.z);
    const float scx = __uint_as_float(w0.w), scy = __uint_as_float(w1.x), scz = __uint_as_float(w1.y);
    // bytes: qlo[3][8] = w1.z w1.w | w2.x w2.y | w2.z w2.w ; qhi[3][8] = w3.x w3.y | w3.z w3.w | w4.x w4.y
    const unsigned ql[6] = {w1.z, w1.w, w2.x, w2.y, w2.z, w2.w}, qh[6] = {w3.x, w3.y, w3.z, w3.w, w4.x, w4.y};
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int wi = c >> 2, sh = 8 * (c & 3);
        const float lx = fmaf((float)((ql[0 + wi] >> sh) & 0xffu), scx, lox), hx = fmaf((float)((qh[0 + wi] >> sh) & 0xffu), scx, lox);
        const float ly = fmaf((float)((ql[2 + wi] >> sh) & 0xffu), scy, loy), hy = fmaf((float)((qh[2 + wi] >> sh) & 0xffu), scy, loy);
        const float lz = fmaf((float)((ql[4 + wi] >> sh) & 0xffu), scz, loz), hz = fmaf((float)((qh[4 + wi] >> sh) & 0xffu), scz, loz);
        const float x0 = fmaf(lx, r.ix, -r.px), x1 = fmaf(hx, r.ix, -r.px);
        const float y0 = fmaf(ly, r.iy, -r.py), y1 = fmaf(hy, r.iy, -r.py);
        const float z0 = fmaf(lz, r.iz, -r.pz), z1 = fmaf(hz, r.iz, -r.pz);
        const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
        const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), LRT_TMAX));
        if (tn <= tf) m |= 1u << c;
    }
    return m;
}

// A candidate joins its ray's bin; its depth-error bound joins the ray's maximum. A bound at the cap (a surfel seen edge-on:
// n.d ~ 0, the depth from the origin says nothing about the depth a re-based round will compute) makes the candidate WILD: its
// key gets t = 0, so it sorts to the front of the bin, and every round of the ray tests the ray's wild candidates first, whatever
// the window says, before it scans the sorted rest. (Giving such rays an infinite margin made a few hundred rays per frame re-scan
// their whole bin every round — 0.14 ms of tail in pass A; handing them to the per-ray fallback cost 1.4 ms: rays sliding along
// facades are common in a street scene.)
__device__ __forceinline__ void wf_append(const WfBufs& w, int ray, float t, int g, float e)
{
    const bool wild = e >= LRT_ERR_CAP;
    const int pos = atomicAdd(w.hit_count + ray, 1) & (WF_TAINT - 1);
    const unsigned long long key = ((unsigned long long)(wild ? 0u : __float_as_uint(t)) << 32) | (unsigned)g;
    if (pos < w.hcap) w.bins[(size_t)ray * w.hcap + pos] = key;
    else if (w.ov_pairs) {                               // beyond the bin: the overflow list; a dropped pair taints the ray (fallback)
        const int j = atomicAdd(w.counts + 14, 1);
        if (j < w.ov_pair_cap) w.ov_pairs[j] = make_uint4((unsigned)ray, 0u, (unsigned)(key & 0xffffffffull), (unsigned)(key >> 32));
        else atomicOr(w.hit_count + ray, WF_TAINT);
    }
    if (wild) atomicAdd(w.nwild + ray, 1);
    else if (e > LRT_ERR_FLOOR) atomicMax(w.emax + ray, __float_as_int(e));
}

// One (ray, leaf) item per thread: the leaf's 8 surfel boxes, then the exact quad test (same arithmetic
// as the per-ray kernels, bounds relaxed: the bin holds CANDIDATES), appended to the ray's bin.
__global__ void __launch_bounds__(256) k_wf_leaf(BvhView bvh, FwdArgs a, WfBufs w, const uint2* __restrict__ in, const int* __restrict__ in_count)
{
    // Two phases per warp-batch of 32 items, so that the expensive part runs with full lanes:
    //   A. every lane evaluates the 8 byte boxes of ITS leaf (uniform code) and drops its (ray, surfel) pairs into a
    //      per-warp shared-memory queue (prefix sum of the pair counts);
    //   B. the queue is consumed 32 pairs at a time: one exact-ish candidate test per lane. Without the queue a lane
    //      loops over its own 0..8 pairs while the others idle: ncu showed 7 active lanes per instruction.
    __shared__ uint2 s_q[8][256];
    const unsigned FULL = 0xffffffffu;
    const int n_in = in ? min(*in_count, w.cap_items) : num_slots(a.R, a.grid_w);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint2* q = s_q[wib];
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n_in; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        int ray = -1; unsigned node = 0;
        if (i < n_in) {
            if (in) { const uint2 it = in[i]; ray = (int)it.x; node = it.y; }
            else ray = slot_to_ray(i, a.R, a.grid_w);
        }
        unsigned m = 0;
        if (ray >= 0) m = leafq_eval(bvh.leafq + node, w.rs[ray]);
        const int cnt = __popc(m);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total == 0) continue;
        int off = incl - cnt;
        while (m) { const int c = __ffs(m) - 1; m &= m - 1; q[off++] = make_uint2((unsigned)ray, node * 8u + c); }
        __syncwarp(FULL);
        for (int j = lane; j < total; j += 32) {
            const uint2 pr = q[j];
            const RaySetup rs = w.rs[pr.x];
            float t, e; int g;
            if (quad_candidate(bvh.rec, (int)pr.y, rs, t, g, e)) wf_append(w, (int)pr.x, t, g, e);
        }
        __syncwarp(FULL);
    }
}

// warp-wide bitonic sort of one key per lane (ascending); T = 64-bit (t bits, id) keys or 32-bit truncated keys
template <typename T>
__device__ __forceinline__ T warp_sort32(T k, int lane)
{
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const T other = __shfl_xor_sync(0xffffffffu, k, stride);
            const bool up = ((lane & size) == 0);
            const bool lower = ((lane & stride) == 0);
            const bool take_min = (up == lower);
            k = take_min ? (k < other ? k : other) : (k < other ? other : k);
        }
    }
    return k;
}

// warp-wide bitonic sort of 64 keys, two per lane: k0 = element `lane`, k1 = element `lane + 32` (ascending)
template <typename T>
__device__ __forceinline__ void warp_sort64(T& k0, T& k1, int lane)
{
#pragma unroll
    for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride == 32) {                            // partner is the lane's own second key; elements 0..63 all sort upwards here
                const T lo = k0 < k1 ? k0 : k1, hi = k0 < k1 ? k1 : k0;
                k0 = lo; k1 = hi;
            } else {
                const T o0 = __shfl_xor_sync(0xffffffffu, k0, stride), o1 = __shfl_xor_sync(0xffffffffu, k1, stride);
                const bool lower = ((lane & stride) == 0);
                const bool up0 = ((lane & size) == 0), up1 = (((lane + 32) & size) == 0);
                k0 = (lower == up0) ? (k0 < o0 ? k0 : o0) : (k0 < o0 ? o0 : k0);
                k1 = (lower == up1) ? (k1 < o1 ? k1 : o1) : (k1 < o1 ? o1 : k1);
            }
        }
    }
}

// One ray, walked by one WARP over its SORTED candidate keys (shared or global memory): the reference's rounds with 32 candidates
// re-tested per step, the round's 16 slots shaded in parallel (16 lanes) and folded in order with shuffles. Used by k_wf_shade (every
// ray) and by k_sp_big (split passes: the rays with more than WF_HCAP candidates — a thread walking thousands of candidates alone
// takes milliseconds, the per-ray fallback of round 1 took 3-4 ms for such a ray).
__device__ __forceinline__ void wf_shade_ray(const unsigned long long* keys, int n, int r, float em, const BvhView& bvh, const FwdArgs& a, int lane)
{
    const unsigned FULL = 0xffffffffu;
    // ---- the reference's round loop (forward.cu:195-292)
    FwdRay q;
    fwd_ray_init(q, r, a);
    for (int round = 0;; round++) {
        unsigned long long slot_key = LRT_KEY_EMPTY;           // lane i < 16 holds slot i of this round
        int nvalid;
        {
            // Candidates are re-tested, exactly, from o' = o + base d, 32 at a time in bin order starting at
            // the first one with t >= base - margin; the 32 best (t', id) keys are kept (bitonic merge). The
            // scan stops when the bin is exhausted or when the next unexamined candidate lies safely beyond
            // the 16th best — so the round's slots are exactly the 16 nearest hits of the re-based ray.
            RaySetup rs;
            ray_setup(rs, q.o, q.d, q.base);
            const float thr = q.base - 2.0f * wf_margin(q.base, em);
            int below = 0;                                      // candidates in front of thr: lower bound in the sorted keys (warp-uniform)
            {
                int lo = 0, hi = n;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__uint_as_float((unsigned)(keys[mid] >> 32)) < thr) lo = mid + 1; else hi = mid;
                }
                below = lo;
            }
            int pos = below;
            unsigned long long best = LRT_KEY_EMPTY;
            // Fast path: candidates are stored in ascending original depth, and re-testing from a point on
            // the same ray preserves that order except between near-ties. So append the valid re-tested keys
            // in bin order (ballot compaction) and only verify that they came out sorted; the general
            // sort + bitonic-merge path below runs only if that check fails.
            int have = 0;
            bool sorted_ok = true;
            const int pos0 = pos;
            for (;;) {
                unsigned long long nk = LRT_KEY_EMPTY;
                const int idx = pos + lane;
                if (idx < n) {
                    const int g = (int)(unsigned)(keys[idx] & 0xffffffffull);
                    float t; int g2;
                    if (quad_hit(bvh.rec_g, g, rs, t, g2)) nk = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)g;
                }
                const unsigned vm = __ballot_sync(FULL, nk != LRT_KEY_EMPTY);
                const int want = lane - have;                                   // lane takes the want-th valid key of this window
                const int src = (want >= 0 && want < __popc(vm)) ? (int)__fns(vm, 0, want + 1) : -1;
                const unsigned long long got = __shfl_sync(FULL, nk, src < 0 ? 0 : src);
                if (src >= 0) best = got;
                have = min(32, have + __popc(vm));
                pos += 32;
                if (pos >= n || have >= 32) break;
                if (have >= LRT_KBUF) {
                    const unsigned long long k16 = __shfl_sync(FULL, best, LRT_KBUF - 1);
                    const float t16 = __uint_as_float((unsigned)(k16 >> 32)) + q.base;
                    const float t_next = __uint_as_float((unsigned)(keys[pos] >> 32));
                    if (t_next - t16 > wf_margin(t16, em)) break;
                }
            }
            {
                const unsigned long long prev = __shfl_up_sync(FULL, best, 1);
                const bool bad = lane > 0 && lane < have && prev > best;
                sorted_ok = __ballot_sync(FULL, bad) == 0;
                // the scan may have stopped at 32 collected keys with candidates left: then the 16th must still be
                // safely in front of the first unexamined candidate
                if (sorted_ok && pos < n) {
                    const unsigned long long k16 = __shfl_sync(FULL, best, LRT_KBUF - 1);
                    const float t16 = __uint_as_float((unsigned)(k16 >> 32)) + q.base;
                    const float t_next = __uint_as_float((unsigned)(keys[pos] >> 32));
                    if (!(have >= LRT_KBUF && t_next - t16 > wf_margin(t16, em))) sorted_ok = false;
                }
                if (sorted_ok && have >= 32) {                                  // valid keys beyond the 32nd were dropped: they must
                    const unsigned long long k16 = __shfl_sync(FULL, best, LRT_KBUF - 1), k32 = __shfl_sync(FULL, best, 31);   // lie safely behind the 16th
                    const float t16 = __uint_as_float((unsigned)(k16 >> 32)), t32 = __uint_as_float((unsigned)(k32 >> 32));
                    if (!(t32 - t16 > wf_margin(t16 + q.base, em))) sorted_ok = false;
                }
            }
            if (!sorted_ok) {                                                   // general path (near-ties re-ordered by re-basing)
            pos = pos0; best = LRT_KEY_EMPTY;
            for (;;) {
                unsigned long long nk = LRT_KEY_EMPTY;
                const int idx = pos + lane;
                if (idx < n) {
                    const int g = (int)(unsigned)(keys[idx] & 0xffffffffull);
                    float t; int g2;
                    if (quad_hit(bvh.rec_g, g, rs, t, g2)) nk = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)g;
                }
                nk = warp_sort32(nk, lane);
                const unsigned long long rev = __shfl_sync(FULL, nk, 31 - lane);      // bitonic merge: 32 smallest of best U nk
                best = warp_sort32(best < rev ? best : rev, lane);
                pos += 32;
                if (pos >= n) break;                                                   // bin exhausted: exact
                const unsigned long long k16 = __shfl_sync(FULL, best, LRT_KBUF - 1);
                if (k16 != LRT_KEY_EMPTY) {
                    const float t16 = __uint_as_float((unsigned)(k16 >> 32)) + q.base;
                    const float t_next = __uint_as_float((unsigned)(keys[pos] >> 32));
                    if (t_next - t16 > wf_margin(t16, em)) break;   // nothing further can rank in the first 16
                }
            }
            }
            slot_key = best;
            nvalid = __popc(__ballot_sync(FULL, best != LRT_KEY_EMPTY));
            // nvalid == 32 only says ">= 32": the round logic below only distinguishes < 16 from >= 16
        }
        const int nr = nvalid < LRT_KBUF ? nvalid : LRT_KBUF;   // slots of this round
        G8Slot sl;
        g8_shade_slot(slot_key, lane < nr, q, bvh, a, sl);
        bool terminated = false;
        for (int i = 0; i < nr; i++) {                          // in-order fold (forward.cu:201-280)
            const float dpt_i = __shfl_sync(FULL, sl.dpt, i);
            const unsigned fl_i = __shfl_sync(FULL, sl.flags, i);
            q.nslots++;
            q.dpt = dpt_i;
            if (!(fl_i & G8_F_DPT_OK)) continue;
            const int g_i = __shfl_sync(FULL, sl.g, i);
            if (g_i == q.last) continue;
            q.last = g_i;
            if (!(fl_i & G8_F_OK)) continue;
            const float alpha = __shfl_sync(FULL, sl.alpha, i);
            q.testT = q.T * (1.0f - alpha);
            if (q.testT < LRT_T_MIN) { terminated = true; break; }
            const float wgt = alpha * q.T;
            const float c0 = __shfl_sync(FULL, sl.c0, i), c1 = __shfl_sync(FULL, sl.c1, i), c2 = __shfl_sync(FULL, sl.c2, i);
            q.C0 += wgt * c0; q.C1 += wgt * c1; q.C2 += wgt * c2;
            q.Dp += wgt * dpt_i; q.W += wgt;
            if (lane == i) {                                     // the owner of the slot commits it (forward.cu:272 + hit list);
                atomicAdd(a.accum_w + g_i, wgt);                  // a ray is only handed to the fallback BEFORE its first round
                if (a.hit_gidx != nullptr && q.ncontrib < a.cap) {
                    a.hit_gidx[(size_t)q.ncontrib * a.R + q.r] = g_i;
                    a.hit_t[(size_t)q.ncontrib * a.R + q.r] = dpt_i;
                    if (a.hit_aux) a.hit_aux[(size_t)q.ncontrib * a.R + q.r] = make_float4(alpha, c0, c1, c2);
                }
            }
            q.ncontrib++;
            q.T = q.testT;
        }
        if (terminated || q.testT < LRT_T_MIN || nvalid < LRT_KBUF) break;      // forward.cu:282-285
        q.base = (float)((double)q.dpt + LRT_STEP_EPS);                          // :288
    }
    if (lane == 0) fwd_write(q, a, 0);
}

// One warp per ray. smem: 4 warps x WF_HCAP keys (16 KB).
#ifndef LRT_SHADE_MIN_BLOCKS
#define LRT_SHADE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(128, LRT_SHADE_MIN_BLOCKS) k_wf_shade(BvhView bvh, FwdArgs a, WfBufs w)
{
    __shared__ unsigned long long s_keys[4][WF_HCAP];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned long long* keys = s_keys[wib];
    const int S = num_slots(a.R, a.grid_w);
    for (int s = blockIdx.x * 4 + wib; s < S; s += gridDim.x * 4) {
        const int r = slot_to_ray(s, a.R, a.grid_w);
        if (r < 0) continue;                                       // warp-uniform
        const int hc = w.hit_count[r];
        if ((hc & WF_TAINT) || hc > WF_HCAP || hc > w.hcap) {        // not representable here: per-ray fallback
            if (lane == 0) w.fb_list[atomicAdd(w.counts + 8, 1)] = r;
            continue;
        }
        const int n = hc;
        const float em = w.nwild[r] > 0 ? __int_as_float(0x7f800000) : __int_as_float(w.emax[r]);      // wild candidates: these kernels scan the whole bin
        // ---- load + sort the bin (bitonic over the next power of two, in shared memory)
        int m = 32; while (m < n) m <<= 1;
        for (int i = lane; i < m; i += 32) keys[i] = i < n ? w.bins[(size_t)r * w.hcap + i] : LRT_KEY_EMPTY;
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
        wf_shade_ray(keys, n, r, em, bvh, a, lane);
        __syncwarp(FULL);
    }
}

// warp-wide bitonic sort of 32 K keys held in registers, K per lane: k[j] = element lane + 32 j (ascending). Partners at
// strides >= 32 are the lane's own keys (no shuffle); with the loops unrolled every direction that does not depend on the lane is
// a compile-time constant.
template <int K>
__device__ __forceinline__ void warp_sort_regs(unsigned long long (&k)[K], int lane)
{
#pragma unroll
    for (int size = 2; size <= 32 * K; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int sj = stride >> 5;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    if ((j & sj) == 0) {
                        const bool up = (((32 * j) & size) == 0);          // size >= 64 here: the bit belongs to j
                        const unsigned long long a = k[j], b = k[j | sj];
                        const unsigned long long lo = a < b ? a : b, hi = a < b ? b : a;
                        k[j] = up ? lo : hi; k[j | sj] = up ? hi : lo;
                    }
                }
            } else {
                const bool lower = ((lane & stride) == 0);
#pragma unroll
                for (int j = 0; j < K; j++) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k[j], stride);
                    const bool up = (((lane + 32 * j) & size) == 0);
                    k[j] = (lower == up) ? (k[j] < o ? k[j] : o) : (k[j] < o ? o : k[j]);
                }
            }
        }
    }
}

// ---- split compositing (default): k_wf_sort (one warp per ray: bitonic sort of the bin, written back) followed by
// k_wf_composite (one THREAD per ray). With traversal gone, what is left per ray is a walk over a short sorted
// candidate list — 16 lanes of a warp shading while 32 replicate the fold (k_wf_shade) costs ~3x the instructions
// of simply letting each lane walk its own ray's list.
__global__ void __launch_bounds__(128) k_wf_sort(FwdArgs a, WfBufs w)
{
    __shared__ unsigned long long s_keys[4][WF_HCAP];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned long long* keys = s_keys[wib];
    for (int r = blockIdx.x * 4 + wib; r < a.R; r += gridDim.x * 4) {
        const int hc = w.hit_count[r];
        if ((hc & WF_TAINT) || hc > w.hcap || hc <= 1) continue;
        const int n = hc;
        unsigned long long* bin = w.bins + (size_t)r * w.hcap;
        if (n <= 32) {                                             // one key per lane: register sort
            unsigned long long k = lane < n ? bin[lane] : LRT_KEY_EMPTY;
            k = warp_sort32(k, lane);
            if (lane < n) bin[lane] = k;
            continue;
        }
        if (n <= 64) {                                             // two keys per lane: still registers only
            unsigned long long k0 = bin[lane], k1 = lane + 32 < n ? bin[lane + 32] : LRT_KEY_EMPTY;
            warp_sort64(k0, k1, lane);
            bin[lane] = k0;
            if (lane + 32 < n) bin[lane + 32] = k1;
            continue;
        }
        if (n <= 128) {                                            // four keys per lane
            unsigned long long k[4];
#pragma unroll
            for (int j = 0; j < 4; j++) k[j] = lane + 32 * j < n ? bin[lane + 32 * j] : LRT_KEY_EMPTY;
            warp_sort_regs<4>(k, lane);
#pragma unroll
            for (int j = 0; j < 4; j++) if (lane + 32 * j < n) bin[lane + 32 * j] = k[j];
            continue;
        }
        if (n <= 256) {                                            // eight keys per lane
            unsigned long long k[8];
#pragma unroll
            for (int j = 0; j < 8; j++) k[j] = lane + 32 * j < n ? bin[lane + 32 * j] : LRT_KEY_EMPTY;
            warp_sort_regs<8>(k, lane);
#pragma unroll
            for (int j = 0; j < 8; j++) if (lane + 32 * j < n) bin[lane + 32 * j] = k[j];
            continue;
        }
        if (n > WF_HCAP) {                                         // a whole block sorts it (k_wf_sort_big): one warp would take ~1 ms alone
            if (lane == 0) w.big_list[atomicAdd(w.counts + 10, 1)] = r;
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
        for (int i = lane; i < n; i += 32) bin[i] = keys[i];
        __syncwarp(FULL);
    }
}

// Bins beyond WF_HCAP candidates (a ray skimming a wall or a vehicle's side: hundreds to thousands), listed by k_wf_sort: one
// 256-thread block per bin, bitonic sort of up to hcap (<= 8192) keys in dynamic shared memory (8 B x hcap).
__global__ void __launch_bounds__(256) k_wf_sort_big(FwdArgs a, WfBufs w, int smem_keys)
{
    extern __shared__ unsigned long long s_big[];
    const int nbig = min(w.counts[10], a.R);
    for (int b = blockIdx.x; b < nbig; b += gridDim.x) {
        const int r = w.big_list[b];
        const int hc = w.hit_count[r] & (WF_TAINT - 1);
        const bool ov = hc > w.hcap;                                // keys in the overflow arena (area padded to a power of two)
        const int n = ov ? hc : min(hc, w.hcap);
        unsigned long long* bin = wf_big_keys(w, r, hc);
        int m = 2 * WF_HCAP; while (m < n) m <<= 1;
        // the block sorts in shared memory what fits there; a longer list (only ever an overflow area, which has its m slots)
        // in place in global memory, one barrier per stage
        unsigned long long* k = m <= smem_keys ? s_big : bin;
        if (k == s_big) { for (int i = threadIdx.x; i < m; i += blockDim.x) s_big[i] = i < n ? bin[i] : LRT_KEY_EMPTY; }
        else { for (int i = n + threadIdx.x; i < m; i += blockDim.x) bin[i] = LRT_KEY_EMPTY; }
        __syncthreads();
        for (int size = 2; size <= m; size <<= 1) {
            for (int ls = 31 - __clz(size) - 1; ls >= 0; ls--) {
                const int stride = 1 << ls;
                for (int i = threadIdx.x; i < (m >> 1); i += blockDim.x) {
                    const int lo = ((i >> ls) << (ls + 1)) + (i & (stride - 1)), hi = lo + stride;
                    const unsigned long long x = k[lo], y = k[hi];
                    const bool up = ((lo & size) == 0);
                    if ((x > y) == up) { k[lo] = y; k[hi] = x; }
                }
                __syncthreads();
            }
        }
        if (k == s_big) for (int i = threadIdx.x; i < n; i += blockDim.x) bin[i] = s_big[i];
        __syncthreads();
    }
}

// (ray, key) pairs of the overflow list -> behind the copy of the ray's bin in its area
__global__ void __launch_bounds__(256) k_ov_scatter(WfBufs w)
{
    const int n = min(w.counts[14], w.ov_pair_cap);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint4 pr = w.ov_pairs[j];
        const int r = (int)pr.x;
        if (w.hit_count[r] & WF_TAINT) continue;
        const int base = w.ov_base[r];
        if (base < 0) continue;                                    // no room in the arena: the ray took the fallback
        const int at = w.hcap + atomicAdd(w.ov_fill + r, 1);
        w.ov_area[(size_t)base + at] = ((unsigned long long)pr.w << 32) | pr.z;
    }
}

// A ray with more candidates than its bin holds (one warp, all lanes): claim an area of the arena, copy the bin there and put the
// ray on the long-bin list. False: no overflow handling here, or the arena is full -> the caller hands the ray to the fallback.
__device__ __forceinline__ bool wf_claim_area(const WfBufs& w, int r, int hc, int lane)
{
    if (!w.ov_pairs) return false;
    int m = 2 * WF_HCAP; while (m < hc) m <<= 1;
    int base = -1;
    if (lane == 0) {
        base = atomicAdd(w.counts + 15, m);
        if (base < 0 || base > w.ov_area_cap - m) base = -1;
        w.ov_base[r] = base;
        if (base >= 0) w.big_list[atomicAdd(w.counts + 10, 1)] = r;
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base < 0) return false;
    const unsigned long long* bin = w.bins + (size_t)r * w.hcap;
    for (int i = lane; i < w.hcap; i += 32) w.ov_area[(size_t)base + i] = bin[i];
    return true;
}

#ifndef LRT_COMPOSITE_MIN_BLOCKS
#define LRT_COMPOSITE_MIN_BLOCKS 3      // 156 registers, no spills; 4-6 blocks were measured slower (spills in the round loop)
#endif
__global__ void __launch_bounds__(128, LRT_COMPOSITE_MIN_BLOCKS) k_wf_composite(BvhView bvh, FwdArgs a, WfBufs w)
{
    // with `order`, a warp's 32 rays have (nearly) equal candidate counts: the per-ray loops below stay in step
    const int S = w.order ? a.R : num_slots(a.R, a.grid_w);
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const int r = w.order ? w.order[s] : slot_to_ray(s, a.R, a.grid_w);
        if (r < 0) continue;
        const int hc = w.hit_count[r];
        if ((hc & WF_TAINT) || hc > w.hcap) { w.fb_list[atomicAdd(w.counts + 8, 1)] = r; continue; }
        const int n = hc;
        const float em = w.nwild[r] > 0 ? __int_as_float(0x7f800000) : __int_as_float(w.emax[r]);      // wild candidates: these kernels scan the whole bin
        const unsigned long long* __restrict__ bin = w.bins + (size_t)r * w.hcap;       // sorted by (t from o, id)
        FwdRay q;
        fwd_ray_init(q, r, a);
        int pos = 0;                                               // first candidate that can still matter
        for (;;) {
            // the round's k-buffer: exact re-test of the candidates from o' = o + base d, in bin order, kept as the
            // (at most) 16 smallest (t', id); stop once the next candidate lies safely behind the 16th
            RaySetup rs;
            ray_setup(rs, q.o, q.d, q.base);
            const float thr = q.base - 2.0f * wf_margin(q.base, em);
            while (pos < n && __uint_as_float((unsigned)(bin[pos] >> 32)) < thr) pos++;
            unsigned long long kb[LRT_KBUF];
            int cnt = 0;
            unsigned long long ck_next = pos < n ? bin[pos] : 0ull;
            for (int i = pos; i < n; i++) {
                const unsigned long long ck = ck_next;
                if (i + 1 < n) ck_next = bin[i + 1];                                       // one ahead: its latency hides behind this test
                if (cnt == LRT_KBUF) {
                    const float t16 = __uint_as_float((unsigned)(kb[LRT_KBUF - 1] >> 32)) + q.base;
                    if (__uint_as_float((unsigned)(ck >> 32)) - t16 > wf_margin(t16, em)) break;
                }
                const int g = (int)(unsigned)(ck & 0xffffffffull);
                float t; int g2;
                if (!quad_hit(bvh.rec_g, g, rs, t, g2)) continue;
                if (a.shs) {   // a slot of this round: most slots composite, so start pulling its SH row (two 128 B lines at D = 3) into L2 now
                    const char* row = reinterpret_cast<const char*>(a.shs + (size_t)g * a.M * 3);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                    if (a.D >= 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
                }
                const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)g;
                if (cnt == LRT_KBUF) { if (key >= kb[LRT_KBUF - 1]) continue; cnt--; }      // replaces the current 16th
                int j = cnt;
                while (j > 0 && kb[j - 1] > key) { kb[j] = kb[j - 1]; j--; }                 // near-sorted input: usually 0 shifts
                kb[j] = key;
                cnt++;
            }
            if (!fwd_shade_round(q, kb, cnt, bvh, a)) break;
        }
        fwd_write(q, a, 0);
    }
}

// k_wf_composite, second organisation (default, LRT_OPT_WAVEFRONT_SHADE = 2): the same walk, arranged so that a ray's thread
// waits for memory as rarely as possible — the kernel is bound by exposed load latency (ncu: >90 % of the stall samples are
// long-scoreboard), not by bandwidth or issue.
//  * candidates are taken four at a time: their keys, then the outer parts of their four records (centre, cutoff, normal), are
//    requested together, so the thread waits once per four gathers instead of once per gather;
//  * a candidate that becomes a slot gets its blending opacity computed right there, from the record it has in registers,
//    with the arithmetic fwd_shade_round() uses (depth = t' + base, point on the ORIGINAL ray): the round's compositing
//    loop then needs no record at all (it used to gather all four 16-byte parts again, usually from L2);
//  * the round's slots (key, opacity) live in shared memory, one column per thread, instead of local memory that competes
//    with the gathers for L1; slots arrive almost sorted, so appending rarely reads them;
//  * the SH basis is evaluated once per ray, and the twelve loads of an SH row are issued before the first use.
// Results are bit-identical to the first organisation (tests/test_gpu_parity.py::test_all_forward_kernels_and_options_agree_bitwise).
#ifndef LRT_COMPOSITE2_MIN_BLOCKS
#define LRT_COMPOSITE2_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(128, LRT_COMPOSITE2_MIN_BLOCKS) k_wf_composite2(BvhView bvh, FwdArgs a, WfBufs w)
{
    __shared__ unsigned long long s_kb[LRT_KBUF][128];             // the round's slots, ascending (t', id)
    __shared__ float s_al[LRT_KBUF][128];                          // their blending opacities (0: cannot contribute)
    const int tx = threadIdx.x;
    const int S = w.order ? a.R : num_slots(a.R, a.grid_w);
    const bool sh_fast = sh_rows_aligned(a);
    const SurfelRec* __restrict__ rec = bvh.rec_g;
    for (int s = blockIdx.x * blockDim.x + tx; s < S; s += gridDim.x * blockDim.x) {
        const int r = w.order ? w.order[s] : slot_to_ray(s, a.R, a.grid_w);
        if (r < 0) continue;
        const int hc = w.hit_count[r];
        if ((hc & WF_TAINT) || hc > w.hcap) { w.fb_list[atomicAdd(w.counts + 8, 1)] = r; continue; }
        const int n = hc;
        const float em = w.nwild[r] > 0 ? __int_as_float(0x7f800000) : __int_as_float(w.emax[r]);      // wild candidates: these kernels scan the whole bin
        const unsigned long long* __restrict__ bin = w.bins + (size_t)r * w.hcap;       // sorted by (t from o, id)
        FwdRay q;
        fwd_ray_init(q, r, a);
        float sb[16];
        const int nb = sh_basis(a.D, q.dirn, sb);
        int pos = 0;                                               // first candidate that can still matter
        int i_end = 0;                                             // where the previous round's scan stopped: everything from there on lies beyond thr
        for (;;) {
            RaySetup rs;
            ray_setup(rs, q.o, q.d, q.base);
            const float thr = q.base - 2.0f * wf_margin(q.base, em);
            {   // first candidate at or beyond thr: lower bound in [pos, i_end] (the bin is sorted by t) — a few dependent loads instead of one per skipped candidate
                int lo = pos, hi = i_end;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__uint_as_float((unsigned)(bin[mid] >> 32)) < thr) lo = mid + 1; else hi = mid;
                }
                pos = lo;
            }
            i_end = n;
            int cnt = 0;
            unsigned long long klast = 0ull;                       // s_kb[cnt - 1]
            bool done = false;
            for (int i0 = pos; i0 < n && !done; i0 += 4) {
                unsigned long long ck4[4];
                float4 b0[4], b3[4];
#pragma unroll
                for (int k = 0; k < 4; k++) ck4[k] = i0 + k < n ? bin[i0 + k] : LRT_KEY_EMPTY;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (i0 + k < n) {
                        const int gk = (int)(unsigned)(ck4[k] & 0xffffffffull);
                        b0[k] = ld_f4(&rec[gk].r0); b3[k] = ld_f4(&rec[gk].r3);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (i0 + k >= n || done) continue;
                    const unsigned long long ck = ck4[k];
                    const float4 a0 = b0[k], a3 = b3[k];
                    if (cnt == LRT_KBUF) {
                        const float t16 = __uint_as_float((unsigned)(klast >> 32)) + q.base;
                        if (__uint_as_float((unsigned)(ck >> 32)) - t16 > wf_margin(t16, em)) { done = true; i_end = i0 + k; continue; }
                    }
                    const int g = (int)(unsigned)(ck & 0xffffffffull);
                    // quad_hit(), operation for operation
                    const float c0 = a0.x - rs.ox, c1 = a0.y - rs.oy, c2 = a0.z - rs.oz;
                    const float den = a3.x * rs.dx + a3.y * rs.dy + a3.z * rs.dz;
                    const float num = a3.x * c0 + a3.y * c1 + a3.z * c2;
                    const float t = num / den;
                    if (!(t > 0.0f)) continue;
                    const float4 a1 = ld_f4(&rec[g].r1), a2 = ld_f4(&rec[g].r2);
                    {
                        const float r0 = (rs.ox + t * rs.dx) - a0.x, r1 = (rs.oy + t * rs.dy) - a0.y, r2 = (rs.oz + t * rs.dz) - a0.z;
                        const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;
                        const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
                        if (!(fabsf(u) <= a0.w && fabsf(v) <= a0.w)) continue;
                        if (!(t < LRT_TMAX)) continue;
                    }
                    const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)g;
                    if (cnt == LRT_KBUF && key >= klast) continue;                              // behind the current 16th
                    if (sh_fast) {   // a slot of this round: most slots composite, so start pulling its SH row (two 128 B lines at D = 3) into L2 now
                        const char* row = reinterpret_cast<const char*>(a.shs + (size_t)g * a.M * 3);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                        if (a.D >= 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
                    }
                    // its opacity, as fwd_shade_round() computes it (forward.cu:212-251)
                    float alpha;
                    {
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
            // the round's compositing (fwd_shade_round()), opacities at hand
            bool terminated = false;
            for (int i = 0; i < cnt; i++) {
                const unsigned long long key = s_kb[i][tx];
                const int g = (int)(unsigned)(key & 0xffffffffull);
                q.nslots++;
                q.dpt = __uint_as_float((unsigned)(key >> 32)) + q.base;                  // forward.cu:212
                if (q.dpt < LRT_MIN_T) continue;                                          // :214
                if (g == q.last) continue;                                                // :220-224
                q.last = g;
                const float alpha = s_al[i][tx];
                if (alpha < 1.0f / 255.0f) continue;
                q.testT = q.T * (1.0f - alpha);
                if (q.testT < LRT_T_MIN) { terminated = true; break; }                    // :253-257
                const float wgt = alpha * q.T;
                float c[3];
                if (sh_fast) {
                    sh_colour_stream_b(nb, sb, a.shs + (size_t)g * a.M * 3, c);
                } else {
                    float sh[48]; bool cl;
                    load_sh_any(a, g, nb, sh);
                    sh_colour<false>(a.D, q.dirn, sh, c, cl, nullptr);
                }
                q.C0 += wgt * c[0]; q.C1 += wgt * c[1]; q.C2 += wgt * c[2];
                q.Dp += wgt * q.dpt; q.W += wgt;
                atomicAdd(a.accum_w + g, wgt);                                            // :272
                if (a.hit_gidx != nullptr && q.ncontrib < a.cap) {
                    a.hit_gidx[(size_t)q.ncontrib * a.R + q.r] = g;
                    a.hit_t[(size_t)q.ncontrib * a.R + q.r] = q.dpt;
                    if (a.hit_aux) a.hit_aux[(size_t)q.ncontrib * a.R + q.r] = make_float4(alpha, c[0], c[1], c[2]);
                }
                q.ncontrib++;
                q.T = q.testT;
            }
            if (terminated || q.testT < LRT_T_MIN || cnt < LRT_KBUF) break;               // :282-285
            q.base = (float)((double)q.dpt + LRT_STEP_EPS);                               // :288
        }
        fwd_write(q, a, 0);
    }
}

// Rays the wavefront handed back (normally none or a handful): the per-ray code path, one thread per ray.
__global__ void __launch_bounds__(128) k_wf_fallback(BvhView bvh, FwdArgs a, WfBufs w)
{
    const int n = min(w.counts[8], a.R);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) forward_one_ray(bvh, a, w.fb_list[i]);
}
// Rays whose contributing-hit list outgrew `cap` in the split passes (lrt_split.cuh): their per-Gaussian weights are already
// accumulated, the colour fold cannot run from a truncated list — the whole ray is redone by the per-ray path with the weight
// accumulation switched off (a.accum_w == nullptr).
__global__ void __launch_bounds__(128) k_wf_fallback_overflow(BvhView bvh, FwdArgs a, WfBufs w)
{
    const int n = min(w.counts[12], a.R);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) forward_one_ray(bvh, a, w.ov_list[i]);
}
