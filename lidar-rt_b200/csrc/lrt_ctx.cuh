// lrt_ctx.cuh — the per-device context behind the C ABI (include/lidar_rt_b200.h).
// Replaces the reference's OptiXState (optix_tracer/optix_wrapper.h:34-49): owns the acceleration
// structure (sorted surfel records + 8-wide implicit hierarchy) and the build workspace.
#pragma once

#include <string>
#include <vector>
#include <map>
#include <cstdio>
#include "../../include/lidar_rt_b200.h"
#include "lrt_common.cuh"

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

#define LRT_INTERNAL_HIT_CAP 256  // hit-list depth of the split forward passes when the caller records no lists
#define LRT_CH_MAX_LEVELS 10      // chamfer point hierarchy, 8-wide: 8^10 points

struct lrt_ctx {
    int device = 0;
    std::string err;
    // acceleration structure
    int P = 0, P_pad = 0, levels = 0;
    int level_off[LRT_MAX_LEVELS] = {0};
    int level_cnt[LRT_MAX_LEVELS] = {0};
    long long n_nodes = 0;
    bool built = false;
    float scale_modifier = 1.0f;
    DevBuf leafq;
    DevBuf wf_ov_pairs, wf_ov_area; // split passes: candidates beyond a bin (overflow list, arena of per-ray areas)
    DevBuf fit_ticket;             // k_fit_top's last-block counter (zero between builds)
    DevBuf rec, nodes, keys_a, keys_b, perm_a, perm_b, rec_g, sort_tmp, bounds, counter;
    DevBuf wf_rs, wf_list_a, wf_list_b, wf_hit_count, wf_bins, wf_fb, wf_ids, wf_keys, wf_sort_tmp, bw_ids, bw_keys, bw_sort_tmp;   // wavefront forward workspace
    DevBuf bw_off, bw_rec_a, bw_rec_b;   // hit-parallel backward (lrt_backward.cu)
    DevBuf dn_pos, dn_tmp;               // densify / prune row compaction (lrt_densify.cu)
    // SH rows read / differentiated in place (lrt_set_sh_parts): host copy of the parts + the device table the kernels search
    DevBuf sh_tab;
    lrt_sh_part sh_parts[LRT_MAX_ASSETS];
    int sh_parts_n = 0, sh_parts_P = 0, sh_parts_M = 0, sh_parts_vec = 0, sh_parts_grad_vec = 0;
    DevBuf sp_cnt, sp_rec, sp_scan_tmp, sp_hits;   // split forward passes (lrt_split.cuh): slice offsets, sorted record stream, internal hit lists
    DevBuf bg_ang, bg_cell_of, bg_cells, bg_sray, bg_wide, bg_plan;   // shared-origin beam grid (lrt_beamgrid.cuh)
    // chamfer distance (lrt_chamfer.cu): one Morton-sorted point hierarchy per cloud, rebuilt every call
    struct ChTree {
        DevBuf pts, boxes;
        int n = 0, n_pad = 0, levels = 0;
        int level_off[LRT_CH_MAX_LEVELS] = {0}, level_cnt[LRT_CH_MAX_LEVELS] = {0};
        size_t bytes() const { return pts.cap + boxes.cap; }
    };
    ChTree ch[2];
    DevBuf ch_tmp, ch_bounds, ch_keys_a, ch_keys_b, ch_idx_a, ch_idx_b;
    // options (lrt_set_option)
    int opt_forward_kernel = 4;   // 0: one thread per ray, 1: persistent threads with per-lane refill, 2: 8 lanes per ray, 3: wavefront,
                                  // 4: shared-origin beam grid (frames with per-ray origins take 3)
    int opt_ray_grid_w = 0;       // > 0: rays are a row-major range image of this width (enables 4 x 8 warp tiles)
    int opt_wavefront_shade = 3;  // wavefront compositing: 0 = one warp per ray (k_wf_shade), 1 = warp-sort + one thread per ray,
                                  // 2 = the same with pipelined record loads and opacities computed when a slot is accepted,
                                  // 3 = split passes over a sorted record stream: slots / colour / fold (lrt_split.cuh, default)
    int opt_beam_cell_pct = 100;  // beam grid: cell edge in percent of the one-ray-per-cell size
    int opt_sort_rays = 1;        // composite / backward replay: process rays in order of descending list length (lanes stay in step)
    int opt_backward_kernel = 2;  // 0: one thread per ray replays its hit list, 1: one warp per ray, one hit per lane (scans),
                                  // 2: two passes (default; needs hit_aux): per-ray prefix pass, then one thread per hit
    int opt_vector_atomics = 1;   // backward: red.global.add.v4.f32 where alignment allows
    int opt_split_fused = 1;      // split passes: 1 = sort + slots in one warp-per-ray kernel (k_sp_warp), 0 = k_sp_sort + record stream + k_sp_slots
    int opt_triangle_depth = 0;   // split passes: hits and depths from the two fp32 proxy triangles (fp64 Moeller-Trumbore) instead of the analytic quad
    int opt_bin_cap = 0;          // > 0: candidate-bin capacity per ray forced to this power of two (tests)
    int opt_sort_key_bits = 16;   // top bits of a 32-bit key the radix sort orders (8 per pass)
    int opt_morton_bits = 32;     // 32: 32-bit cubic-cell keys (default); 63: 21 bits/axis on cubic cells; 30: 10 bits/axis, per-axis extent
    int fwd_blocks_per_sm = 0, g8_blocks_per_sm = 0, num_sms = 0;
    long long builds = 0, refits = 0;
    int launches = 0;

    // optional live per-kernel timing (LRT_OPT_KERNEL_TIMING): CUDA events around every launch, on the launch stream
    int opt_kernel_timing = 0;
    cudaStream_t side_stream = nullptr;     // backward: gradient zero-fill beside the count / scan / prefix passes
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    struct TimedSpan { const char* name; cudaEvent_t a, b; };
    std::vector<TimedSpan> spans;           // spans of the calls since the last lrt_get_kernel_times
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t take_event()
    {
        cudaEvent_t e = nullptr;
        if (!event_pool.empty()) { e = event_pool.back(); event_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    void span_begin(const char* name, cudaStream_t s)
    {
        if (!opt_kernel_timing) return;
        TimedSpan t; t.name = name; t.a = take_event(); t.b = take_event();
        cudaEventRecord(t.a, s);
        spans.push_back(t);
    }
    void span_end(cudaStream_t s)
    {
        if (!opt_kernel_timing || spans.empty()) return;
        cudaEventRecord(spans.back().b, s);
    }

    void set_error(const char* what, cudaError_t e)
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "%s failed: %s", what, cudaGetErrorString(e));
        err = buf;
    }
    void set_error(const char* msg) { err = msg; }

    cudaError_t reserve(DevBuf& b, size_t bytes)
    {
        if (bytes <= b.cap) return cudaSuccess;
        if (b.p) { cudaError_t e = cudaFree(b.p); b.p = nullptr; b.cap = 0; if (e != cudaSuccess) return e; }
        const size_t want = bytes + bytes / 4 + 256;      // headroom: densification grows P every 100 iterations
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e == cudaSuccess) b.cap = want;
        return e;
    }
    size_t total_bytes() const
    {
        return leafq.cap + fit_ticket.cap + wf_ov_pairs.cap + wf_ov_area.cap + rec.cap + nodes.cap + keys_a.cap + keys_b.cap + perm_a.cap + perm_b.cap + rec_g.cap + sort_tmp.cap + bounds.cap + counter.cap + wf_rs.cap + wf_list_a.cap + wf_list_b.cap +
               wf_hit_count.cap + wf_bins.cap + wf_fb.cap + wf_ids.cap + wf_keys.cap + wf_sort_tmp.cap + bw_ids.cap + bw_keys.cap + bw_sort_tmp.cap +
               bg_ang.cap + bg_cell_of.cap + bg_cells.cap + bg_sray.cap + bg_wide.cap + bg_plan.cap +
               bw_off.cap + bw_rec_a.cap + bw_rec_b.cap + sp_cnt.cap + sp_rec.cap + sp_scan_tmp.cap + sp_hits.cap + dn_pos.cap + dn_tmp.cap + sh_tab.cap + ch[0].bytes() + ch[1].bytes() + ch_tmp.cap + ch_bounds.cap + ch_keys_a.cap + ch_keys_b.cap + ch_idx_a.cap + ch_idx_b.cap;
    }
    BvhView view() const
    {
        BvhView v;
        v.rec = (const SurfelRec*)rec.p;
        v.nodes = (const Node8*)nodes.p;
        v.rec_g = (const SurfelRec*)rec_g.p;
        v.leafq = (const LeafQ*)leafq.p;
        for (int i = 0; i < LRT_MAX_LEVELS; i++) v.level_off[i] = level_off[i];
        v.levels = levels;
        v.P = P;
        return v;
    }
};

// implemented in lrt_build.cu / lrt_forward.cu / lrt_backward.cu
int lrt_build_impl(lrt_ctx* ctx, int P, const float* means, const float* scales, const float* rots,
                   const float* opac, float mod, bool refit, cudaStream_t s);
int lrt_forward_impl(lrt_ctx* ctx, int R, const float* ray_o, int ray_o_stride, const float* ray_d,
                     const float* bg, int P, const float* means, const float* scales, const float* rots,
                     const float* opac, const float* shs, int D, int M, float mod,
                     float* out, float* accum_w, int32_t* hit_gidx, float* hit_t, float* hit_aux, int32_t* hit_cnt,
                     int cap, int32_t* slot_cnt, cudaStream_t s);
int lrt_backward_impl(lrt_ctx* ctx, int R, const float* ray_o, int ray_o_stride, const float* ray_d,
                      const float* bg, int P, const float* means, const float* scales, const float* rots,
                      const float* opac, const float* shs, int D, int M, float mod,
                      const float* fwd_out, const float* dL_dout,
                      const int32_t* hit_gidx, const float* hit_t, const float* hit_aux, const int32_t* hit_cnt, int cap,
                      float* dL_dmeans, float* dL_dshs, float* dL_dopac, float* dL_dscales,
                      float* dL_drots, int flags, cudaStream_t s);
int lrt_prepare_impl(lrt_ctx* ctx, int n_assets, const lrt_asset* assets, int M, float* means, float* scales, float* rots,
                     float* opac, float* shs, cudaStream_t s);
int lrt_prepare_backward_impl(lrt_ctx* ctx, int n_assets, const lrt_asset* assets, int M, const float* g_means, const float* g_scales,
                              const float* g_rots, const float* g_opac, const float* g_shs, cudaStream_t s);
int lrt_chamfer_forward_impl(lrt_ctx* ctx, int b, int n, const float* xyz1, int m, const float* xyz2,
                             float* dist1, int32_t* idx1, float* dist2, int32_t* idx2, cudaStream_t s);
int lrt_chamfer_backward_impl(lrt_ctx* ctx, int b, int n, const float* xyz1, int m, const float* xyz2,
                              const float* grad_dist1, const float* grad_dist2, const int32_t* idx1, const int32_t* idx2,
                              float* grad_xyz1, float* grad_xyz2, cudaStream_t s);
int lrt_adam_step_impl(lrt_ctx* ctx, int n_tensors, const lrt_adam_tensor* tensors, double beta1, double beta2, double eps, cudaStream_t s);
int lrt_compact_rows_impl(lrt_ctx* ctx, int n_rows, const unsigned char* keep, int n_tensors, const lrt_row_tensor* tensors, cudaStream_t s);
int lrt_densify_rows_impl(lrt_ctx* ctx, int P, const unsigned char* clone_mask, const unsigned char* split_mask, int n_clone, int n_split,
                          int N, const float* samples, const float* rotation, int n_tensors, const lrt_row_tensor* tensors, cudaStream_t s);
int lrt_range_rays_impl(lrt_ctx* ctx, int H, int W, const float* inc_table, float inc_lo, float inc_hi, float pixel_offset,
                        float angle_offset, const float* sensor2world, const float* range_map, float* out, float* centre, cudaStream_t s);
