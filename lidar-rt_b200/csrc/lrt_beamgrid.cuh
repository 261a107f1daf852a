// lrt_beamgrid.cuh — forward "kernel E": candidate binning for rays that SHARE ONE ORIGIN (a LiDAR frame).
// Included by lrt_forward.cu inside its anonymous namespace, after lrt_wavefront.cuh (shares WfBufs, the per-ray
// candidate bins, k_wf_sort / k_wf_composite / k_wf_fallback).
//
// LiDARSensor.get_range_rays() returns one sensor centre expanded over (H, W) (lidar_sensor.py:400, the
// stride-0 origin of include/lidar_rt_b200.h). With a common origin, "which rays can hit this surfel" is a
// question about DIRECTIONS only: the surfel's proxy quad subtends a small (azimuth, elevation) window and only
// the rays inside that window need the quad test. So instead of walking the hierarchy once per ray
// (k_wf_level x levels + k_wf_leaf: ~160 node evaluations per ray, 1.1 ms per frame), the frame's rays are
// dropped into a uniform (azimuth, elevation) grid (counting sort, 170 k items) and ONE pass over the surfel
// records does the rest: each surfel computes a conservative window from its centre and half-extent vectors,
// and tests exactly the rays stored in the cells of that window. The candidate test, the bins, the per-ray
// sort and the compositing kernel are the wavefront's, so results are bit-identical to kernels A-D.
//
// No assumption is made about the scan pattern: any set of directions works, the grid only adapts its cell
// size to the rays' elevation span and count. Rays with per-ray origins (stride 3) take the wavefront path.
#pragma once

#define BG_PI 3.14159265358979323846f
#define BG_MAX_NA 8192              // azimuth cells (upper bound)
#define BG_HEAVY_CELLS 96           // a surfel whose window covers more cells than this is split into one work item per elevation row
#define BG_ITEM_CAP (1 << 22)       // (surfel, row) items of the heavy pass; beyond it a heavy surfel is simply handled inline
#define BG_ANG_PAD 2e-5f            // radians added to every window edge (atan2f / asinf / floor rounding is < 1e-6)

struct BgPlan { float el_lo, inv_de, inv_da; int NA, NE; };

struct BgBufs {
    float2* ang;                    // (R) azimuth, elevation of every ray direction
    int* el_bounds;                 // [0] max el, [1] max -el as ordered ints (memset 0x80 = -inf)
    BgPlan* plan;
    int* cell_of;                   // (R)
    int* cell_cnt;                  // (ncell_cap + 1) counts, consumed as fill cursors
    int* cell_start;                // (ncell_cap + 1) exclusive scan of cell_cnt
    float4* sray;                   // (R) rays in cell order: (dx, dy, dz, ray id bits)
    int4* items; int* item_count;  // heavy pass: (surfel record, first cell, last cell, -) = one contiguous run of the sorted rays
    int ncell_cap;
};

__device__ __forceinline__ int bg_f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float bg_ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ---- 1. ray directions -> (azimuth, elevation); elevation span of the frame; per-ray bin counters reset
__global__ void __launch_bounds__(256) k_bg_angles(FwdArgs a, WfBufs w, BgBufs b)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float el = 0.0f; bool ok = false;
    if (r < a.R) {
        const float dx = a.ray_d[3 * (size_t)r], dy = a.ray_d[3 * (size_t)r + 1], dz = a.ray_d[3 * (size_t)r + 2];
        const float az = atan2f(dy, dx);
        el = atan2f(dz, sqrtf(dx * dx + dy * dy));
        ok = (az == az) && (el == el);
        b.ang[r] = ok ? make_float2(az, el) : make_float2(0.0f, 0.0f);     // NaN direction: hits nothing anyway
        w.hit_count[r] = 0;
        w.emax[r] = 0; w.nwild[r] = 0;
        w.ray_ids[r] = r;
        if (w.ov_fill) w.ov_fill[r] = 0;
    }
    float hi = ok ? el : -4.0f, nlo = ok ? -el : -4.0f;
#pragma unroll
    for (int o = 16; o; o >>= 1) { hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); nlo = fmaxf(nlo, __shfl_xor_sync(0xffffffffu, nlo, o)); }
    if ((threadIdx.x & 31) == 0 && hi > -3.0f) { atomicMax(b.el_bounds, bg_f2ord(hi)); atomicMax(b.el_bounds + 1, bg_f2ord(nlo)); }
}

// ---- 2. grid dimensions: square cells sized for ~1 ray per cell over the band the rays occupy
__global__ void k_bg_plan(int R, BgBufs b, float cell_scale)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float el_hi = bg_ord2f(b.el_bounds[0]), el_lo = -bg_ord2f(b.el_bounds[1]);
    if (!(el_hi >= el_lo)) { el_hi = 0.0f; el_lo = 0.0f; }                 // no valid ray
    el_hi = fminf(el_hi + 1e-4f, 0.5f * BG_PI); el_lo = fmaxf(el_lo - 1e-4f, -0.5f * BG_PI);
    const float span = fmaxf(el_hi - el_lo, 1e-3f);
    float delta = cell_scale * sqrtf(2.0f * BG_PI * span / (float)max(R, 1));
    delta = fmaxf(delta, 2.0f * BG_PI / (float)BG_MAX_NA);
    int NA = (int)ceilf(2.0f * BG_PI / delta); NA = max(8, min(NA, BG_MAX_NA));
    int NE = (int)ceilf(span / delta); NE = max(1, min(NE, b.ncell_cap / NA));
    BgPlan p;
    p.NA = NA; p.NE = NE; p.el_lo = el_lo; p.inv_de = (float)NE / span; p.inv_da = (float)NA / (2.0f * BG_PI);
    *b.plan = p;
}

__device__ __forceinline__ int bg_cell_a(const BgPlan& p, float az) { return (int)floorf((az + BG_PI) * p.inv_da); }
__device__ __forceinline__ int bg_cell_e(const BgPlan& p, float el) { return (int)floorf((el - p.el_lo) * p.inv_de); }

// ---- 3. counting sort of the rays by cell
__global__ void __launch_bounds__(256) k_bg_count(int R, BgBufs b)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const BgPlan p = *b.plan;
    const float2 ae = b.ang[r];
    const int ia = min(max(bg_cell_a(p, ae.x), 0), p.NA - 1), ie = min(max(bg_cell_e(p, ae.y), 0), p.NE - 1);
    const int c = ie * p.NA + ia;
    b.cell_of[r] = c;
    atomicAdd(b.cell_cnt + c, 1);
}

__global__ void __launch_bounds__(256) k_bg_fill(FwdArgs a, BgBufs b)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const int c = b.cell_of[r];
    const int pos = b.cell_start[c] + atomicSub(b.cell_cnt + c, 1) - 1;
    b.sray[pos] = make_float4(a.ray_d[3 * (size_t)r], a.ray_d[3 * (size_t)r + 1], a.ray_d[3 * (size_t)r + 2], __int_as_float(r));
}

// Conservative (azimuth, elevation) window of one surfel's RELAXED proxy quad (the quad_candidate() bounds) seen
// from `o`. The quad is { c + s a + t b : |s|, |t| <= 1 } with c = mu - o and a, b its half-extent vectors, so
//   z in [c.z - ez, c.z + ez],  ez = |a.z| + |b.z|
//   horizontal position within exy = |a.xy| + |b.xy| of c.xy
// which bounds elevation by the four corner cases of atan2(z, h) and azimuth by asin(exy / |c.xy|). Ground
// surfels seen at grazing angles get a window as thin as they look (a bounding cone would be ~10x taller).
// Returns false if the surfel cannot be hit by any ray of the grid. If the sensor sits inside (or next to) the quad's
// bounding sphere or on its vertical axis, the window spans every azimuth.
struct BgWindow { int ia_lo, na, ie_lo, ne; };

__device__ __forceinline__ bool bg_window(const SurfelRec* __restrict__ rec, int i, const float* o, const BgPlan& p, BgWindow& win)
{
    const float4 r0 = ld_f4(&rec[i].r0);
    const float f = r0.w;
    if (!(f >= 0.0f && f < 1e30f)) return false;                           // invalid proxy (opacity < 1/255: NaN) or padding
    const float4 r1 = ld_f4(&rec[i].r1), r2 = ld_f4(&rec[i].r2);
    const float lim = (f + 1e-3f * (1.0f + f)) * 1.0001f;                 // quad_candidate's bound
    const float su = lim / (r1.x * r1.x + r1.y * r1.y + r1.z * r1.z), sv = lim / (r2.x * r2.x + r2.y * r2.y + r2.z * r2.z);
    const float ax = r1.x * su, ay = r1.y * su, az_ = r1.z * su, bx = r2.x * sv, by = r2.y * sv, bz = r2.z * sv;
    const float cx = r0.x - o[0], cy = r0.y - o[1], cz = r0.z - o[2];
    const float dist = sqrtf(cx * cx + cy * cy + cz * cz);
    if (!(dist < 1e30f) || !(su < 1e30f) || !(sv < 1e30f)) return false;
    float pad = 1e-4f + 2e-5f * dist;                                      // world-space slack, >> fp32 rounding of the quad test
    {   // + what quad_candidate() allows a hit to move in the surfel's plane: its depth-error bound for the least favourable ray
        // that can hit the (relaxed) quad. Points x of the quad's plane have n.x = n.c, so n.d >= |n.c| / (dist + reach).
        const float r3x = __ldg(&rec[i].r3.x), r3y = __ldg(&rec[i].r3.y), r3z = __ldg(&rec[i].r3.z);
        const float reach = sqrtf(ax * ax + ay * ay + az_ * az_) + sqrtf(bx * bx + by * by + bz * bz);
        const float tmax = dist + reach;
        const float den_min = fabsf(r3x * cx + r3y * cy + r3z * cz) / tmax;
        const float Sc = fabsf(r3x) * (fabsf(cx) + reach) + fabsf(r3y) * (fabsf(cy) + reach) + fabsf(r3z) * (fabsf(cz) + reach);
        const float So = fabsf(r3x * o[0]) + fabsf(r3y * o[1]) + fabsf(r3z * o[2]);
        float e = 1.2e-7f * ((8.0f * Sc + 12.0f * tmax + So) / den_min) + 3.6e-7f * tmax;
        if (!(e < LRT_ERR_CAP)) e = LRT_ERR_CAP;
        pad += e;
    }
    const float ez = fabsf(az_) + fabsf(bz) + pad;
    const float exy = sqrtf(ax * ax + ay * ay) + sqrtf(bx * bx + by * by) + pad;
    const float rad = sqrtf(ax * ax + ay * ay + az_ * az_) + sqrtf(bx * bx + by * by + bz * bz) + pad;
    const float rho = sqrtf(cx * cx + cy * cy);
    float el_lo, el_hi; bool all_az;
    if (dist <= rad + 2e-3f) { el_lo = -BG_PI; el_hi = BG_PI; all_az = true; }       // sensor inside the bounding sphere
    else {
        const float zlo = cz - ez, zhi = cz + ez, hlo = fmaxf(rho - exy, 0.0f), hhi = rho + exy;
        el_hi = atan2f(zhi, zhi > 0.0f ? hlo : hhi) + BG_ANG_PAD;
        el_lo = atan2f(zlo, zlo > 0.0f ? hhi : hlo) - BG_ANG_PAD;
        all_az = !(rho > exy * 1.0001f);
    }
    int e0 = bg_cell_e(p, el_lo), e1 = bg_cell_e(p, el_hi);
    if (e1 < 0 || e0 >= p.NE) return false;                                // outside the band the rays occupy
    e0 = max(e0, 0); e1 = min(e1, p.NE - 1);
    win.ie_lo = e0; win.ne = e1 - e0 + 1;
    if (all_az) { win.ia_lo = 0; win.na = p.NA; return true; }
    // azimuth: the quad's horizontal projection is a convex polygon that does not contain the origin (rho > exy), so its
    // azimuth range is exactly the range of its 4 corners; widened by the slack `pad` seen from the nearest possible
    // horizontal distance, and never wider than the bounding-circle window asin(exy / rho)
    const float az0 = atan2f(cy, cx), dlt = asinf(fminf(exy / rho, 1.0f)) + BG_ANG_PAD;
    float lo = 0.0f, hi = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float sa = (k & 1) ? 1.0f : -1.0f, sb = (k & 2) ? 1.0f : -1.0f;
        float dk = atan2f(cy + sa * ay + sb * by, cx + sa * ax + sb * bx) - az0;
        dk = dk > BG_PI ? dk - 2.0f * BG_PI : (dk < -BG_PI ? dk + 2.0f * BG_PI : dk);
        lo = fminf(lo, dk); hi = fmaxf(hi, dk);
    }
    const float apad = asinf(fminf(pad / (rho - exy + pad), 1.0f)) + BG_ANG_PAD;
    lo = fmaxf(lo - apad, -dlt); hi = fminf(hi + apad, dlt);
    const int a0 = bg_cell_a(p, az0 + lo), a1 = bg_cell_a(p, az0 + hi);
    win.na = min(a1 - a0 + 1, p.NA);
    win.ia_lo = ((a0 % p.NA) + p.NA) % p.NA;
    return true;
}

__device__ __forceinline__ void bg_test_and_append(const SurfelRec* __restrict__ rec, int i, const float* o, const float4 sr, const WfBufs& w)
{
    RaySetup rs;
    rs.ox = o[0]; rs.oy = o[1]; rs.oz = o[2]; rs.dx = sr.x; rs.dy = sr.y; rs.dz = sr.z;
    float t, e; int g;
    if (quad_candidate(rec, i, rs, t, g, e)) wf_append(w, __float_as_int(sr.w), t, g, e);
}

// ---- 4. one pass over the surfel records. A warp takes 32 consecutive records (Morton order: neighbours in space,
// so neighbouring windows) and each lane computes its surfel's window. The work is then flattened twice so that
// every lane does exactly one quad test per trip, whatever the window sizes:
//   a. windows -> (surfel, elevation row [, wrap half]) segments; a segment is a CONTIGUOUS run of the cell-sorted ray
//      array, found with two cell_start loads;
//   b. 32 segments at a time -> their rays: prefix sum of the run lengths, and lane k of a trip takes the k-th ray of
//      the concatenated runs (binary search over 32 shared-memory prefixes).
// Consecutive lanes therefore test consecutive rays of the sorted array (coalesced 16 B loads) against the same
// surfel record (one broadcast load). Window sizes go with 1 / distance^2, so the few surfels near the sensor would
// keep their warp busy 100x longer than the rest: windows above BG_HEAVY_CELLS cells are not processed here but
// queued, one item per elevation row, for k_bg_heavy where every item gets a warp of its own.
__global__ void __launch_bounds__(256) k_bg_bin(BvhView bvh, FwdArgs a, WfBufs w, BgBufs b, int n_rec)
{
    __shared__ int4 s_win[8][32];          // per lane: exclusive segment offset, ia_lo, na | (wrap << 30), ie_lo
    __shared__ int s_incl[8][32];          // inclusive segment prefix
    __shared__ int s_rincl[8][32];         // per segment of the current 32: inclusive ray prefix
    __shared__ int2 s_seg[8][32];          // per segment: (first ray position - exclusive ray prefix, owning lane)
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const BgPlan p = *b.plan;
    const float o[3] = {a.ray_o[0], a.ray_o[1], a.ray_o[2]};
    const int* __restrict__ cs = b.cell_start;
    const float4* __restrict__ sray = b.sray;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n_rec; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        BgWindow win; win.ia_lo = 0; win.na = 0; win.ie_lo = 0; win.ne = 0;
        int nseg = 0, wrap = 0;
        if (i < n_rec && bg_window(bvh.rec, i, o, p, win)) {
            bool inline_ = true;
            if (win.na * win.ne > BG_HEAVY_CELLS) {
                const int wr = (win.ia_lo + win.na > p.NA) ? 1 : 0, ni = win.ne << wr;
                const int at = atomicAdd(b.item_count, ni);
                if (at + ni <= BG_ITEM_CAP) {
                    for (int r = 0; r < win.ne; r++) {
                        const int rb = (win.ie_lo + r) * p.NA;
                        if (!wr) b.items[at + r] = make_int4(i, rb + win.ia_lo, rb + win.ia_lo + win.na - 1, 0);
                        else {
                            b.items[at + 2 * r] = make_int4(i, rb + win.ia_lo, rb + p.NA - 1, 0);
                            b.items[at + 2 * r + 1] = make_int4(i, rb, rb + win.ia_lo + win.na - 1 - p.NA, 0);
                        }
                    }
                    inline_ = false;
                } else {
                    for (int r = at; r < BG_ITEM_CAP; r++) b.items[r] = make_int4(-1, 0, 0, 0);   // queue full: void the tail, handle inline
                }
            }
            if (inline_) { wrap = (win.ia_lo + win.na > p.NA) ? 1 : 0; nseg = win.ne << wrap; }
        }
        int incl = nseg;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(FULL, incl, s); if (lane >= s) incl += v; }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total == 0) continue;
        __syncwarp(FULL);
        s_win[wib][lane] = make_int4(incl - nseg, win.ia_lo, win.na | (wrap << 30), win.ie_lo);
        s_incl[wib][lane] = incl;
        __syncwarp(FULL);
        for (int sb = 0; sb < total; sb += 32) {
            const int j = sb + lane;
            int cnt = 0, p0 = 0, L = 0;
            if (j < total) {
#pragma unroll
                for (int s = 16; s; s >>= 1) if (s_incl[wib][L + s - 1] <= j) L += s;            // first lane whose inclusive prefix exceeds j
                const int4 wv = s_win[wib][L];
                const int k = j - wv.x, wr = wv.z >> 30, na = wv.z & 0x3fffffff;
                const int row = wv.w + (k >> wr);
                int c0 = wv.y, c1 = wv.y + na - 1;
                if (wr) { if (k & 1) { c0 = 0; c1 = wv.y + na - 1 - p.NA; } else c1 = p.NA - 1; }
                p0 = __ldg(cs + row * p.NA + c0);
                cnt = __ldg(cs + row * p.NA + c1 + 1) - p0;
            }
            int rincl = cnt;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(FULL, rincl, s); if (lane >= s) rincl += v; }
            const int rtotal = __shfl_sync(FULL, rincl, 31);
            if (rtotal == 0) continue;
            __syncwarp(FULL);
            s_rincl[wib][lane] = rincl;
            s_seg[wib][lane] = make_int2(p0 - (rincl - cnt), L);
            __syncwarp(FULL);
            for (int k = lane; k < rtotal; k += 32) {
                int S = 0;
#pragma unroll
                for (int s = 16; s; s >>= 1) if (s_rincl[wib][S + s - 1] <= k) S += s;
                const int2 sg = s_seg[wib][S];
                bg_test_and_append(bvh.rec, base + sg.y, o, __ldg(sray + sg.x + k), w);
            }
        }
    }
}

// ---- 5. the heavy windows (surfels close to the sensor, or with the sensor inside them / on their vertical axis),
// already cut into contiguous ray runs: a warp takes 32 runs at a time and flattens them exactly like step 4b, so the
// runs of one big surfel are spread over many warps and every lane still does one test per trip.
__global__ void __launch_bounds__(256) k_bg_heavy(BvhView bvh, FwdArgs a, WfBufs w, BgBufs b)
{
    __shared__ int s_rincl[8][32];
    __shared__ int2 s_seg[8][32];
    const unsigned FULL = 0xffffffffu;
    const int n = min(*b.item_count, BG_ITEM_CAP);
    if (blockIdx.x == 0 && threadIdx.x == 0) w.counts[9] = n;              // lrt_debug_counters
    if (n == 0) return;
    const float o[3] = {a.ray_o[0], a.ray_o[1], a.ray_o[2]};
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int* __restrict__ cs = b.cell_start;
    const float4* __restrict__ sray = b.sray;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n; base += gridDim.x * blockDim.x) {
        const int j = base + lane;
        int cnt = 0, p0 = 0, surf = -1;
        if (j < n) {
            const int4 it = b.items[j];
            if (it.x >= 0) { surf = it.x; p0 = __ldg(cs + it.y); cnt = __ldg(cs + it.z + 1) - p0; }
        }
        int rincl = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(FULL, rincl, s); if (lane >= s) rincl += v; }
        const int rtotal = __shfl_sync(FULL, rincl, 31);
        if (rtotal == 0) continue;
        __syncwarp(FULL);
        s_rincl[wib][lane] = rincl;
        s_seg[wib][lane] = make_int2(p0 - (rincl - cnt), surf);
        __syncwarp(FULL);
        for (int k = lane; k < rtotal; k += 32) {
            int S = 0;
#pragma unroll
            for (int s = 16; s; s >>= 1) if (s_rincl[wib][S + s - 1] <= k) S += s;
            const int2 sg = s_seg[wib][S];
            bg_test_and_append(bvh.rec, sg.y, o, __ldg(sray + sg.x + k), w);
        }
    }
}
