// lrt_api.cu — the extern "C" surface declared in include/lidar_rt_b200.h.
#include <new>
#include "lrt_ctx.cuh"

static std::string g_create_error;

#ifdef LRT_STATS
__device__ unsigned long long g_lrt_stats[16];
#endif

extern "C" {

int lrt_version(void) { return LRT_VERSION; }

int lrt_ctx_create(int device, lrt_ctx** out_ctx)
{
    if (!out_ctx) { g_create_error = "lrt_ctx_create: out_ctx is null"; return LRT_ERR_INVALID; }
    *out_ctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_create_error = std::string("lrt_ctx_create: no CUDA device (") + cudaGetErrorString(e) + ")";
        return LRT_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_create_error = "lrt_ctx_create: device index out of range"; return LRT_ERR_INVALID; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice failed: ") + cudaGetErrorString(e); return LRT_ERR_CUDA; }
    lrt_ctx* c = new (std::nothrow) lrt_ctx();
    if (!c) { g_create_error = "lrt_ctx_create: out of host memory"; return LRT_ERR_INVALID; }
    c->device = device;
    *out_ctx = c;
    return LRT_OK;
}

int lrt_ctx_destroy(lrt_ctx* ctx)
{
    if (!ctx) return LRT_OK;
    cudaSetDevice(ctx->device);
    for (auto& t : ctx->spans) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    DevBuf* bufs[] = {&ctx->leafq, &ctx->fit_ticket, &ctx->wf_ov_pairs, &ctx->wf_ov_area, &ctx->rec, &ctx->nodes, &ctx->keys_a, &ctx->keys_b, &ctx->perm_a, &ctx->perm_b, &ctx->rec_g, &ctx->sort_tmp, &ctx->bounds, &ctx->counter,
                      &ctx->wf_rs, &ctx->wf_list_a, &ctx->wf_list_b, &ctx->wf_hit_count, &ctx->wf_bins, &ctx->wf_fb,
                      &ctx->wf_ids, &ctx->wf_keys, &ctx->wf_sort_tmp, &ctx->bw_ids, &ctx->bw_keys, &ctx->bw_sort_tmp,
                      &ctx->bw_off, &ctx->bw_rec_a, &ctx->bw_rec_b, &ctx->dn_pos, &ctx->dn_tmp, &ctx->sh_tab, &ctx->sp_cnt, &ctx->sp_rec, &ctx->sp_scan_tmp, &ctx->sp_hits, &ctx->bg_ang, &ctx->bg_cell_of, &ctx->bg_cells, &ctx->bg_sray, &ctx->bg_wide, &ctx->bg_plan,
                      &ctx->ch_tmp, &ctx->ch_bounds, &ctx->ch_keys_a, &ctx->ch_keys_b, &ctx->ch_idx_a, &ctx->ch_idx_b,
                      &ctx->ch[0].pts, &ctx->ch[0].boxes, &ctx->ch[1].pts, &ctx->ch[1].boxes};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    delete ctx;
    return LRT_OK;
}

const char* lrt_last_error(const lrt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int lrt_build(lrt_ctx* ctx, int P, const float* means, const float* scales, const float* rots,
              const float* opac, float scale_modifier, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_build_impl(ctx, P, means, scales, rots, opac, scale_modifier, false, (cudaStream_t)stream);
}

int lrt_refit(lrt_ctx* ctx, int P, const float* means, const float* scales, const float* rots,
              const float* opac, float scale_modifier, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_build_impl(ctx, P, means, scales, rots, opac, scale_modifier, true, (cudaStream_t)stream);
}

int lrt_forward(lrt_ctx* ctx, int R, const float* ray_o, int ray_o_stride, const float* ray_d,
                const float* bg, int P, const float* means, const float* scales, const float* rots,
                const float* opac, const float* shs, int D, int M, float scale_modifier,
                float* out, float* accum_w, int32_t* hit_gidx, float* hit_t, float* hit_aux, int32_t* hit_cnt,
                int cap, int32_t* slot_cnt, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_forward_impl(ctx, R, ray_o, ray_o_stride, ray_d, bg, P, means, scales, rots, opac, shs, D, M,
                            scale_modifier, out, accum_w, hit_gidx, hit_t, hit_aux, hit_cnt, cap, slot_cnt, (cudaStream_t)stream);
}

int lrt_backward(lrt_ctx* ctx, int R, const float* ray_o, int ray_o_stride, const float* ray_d,
                 const float* bg, int P, const float* means, const float* scales, const float* rots,
                 const float* opac, const float* shs, int D, int M, float scale_modifier,
                 const float* fwd_out, const float* dL_dout,
                 const int32_t* hit_gidx, const float* hit_t, const float* hit_aux, const int32_t* hit_cnt, int cap,
                 float* dL_dmeans, float* dL_dshs, float* dL_dopac, float* dL_dscales,
                 float* dL_drots, int flags, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_backward_impl(ctx, R, ray_o, ray_o_stride, ray_d, bg, P, means, scales, rots, opac, shs, D, M,
                             scale_modifier, fwd_out, dL_dout, hit_gidx, hit_t, hit_aux, hit_cnt, cap,
                             dL_dmeans, dL_dshs, dL_dopac, dL_dscales, dL_drots, flags, (cudaStream_t)stream);
}

int lrt_prepare(lrt_ctx* ctx, int n_assets, const lrt_asset* assets, int M, float* means, float* scales, float* rots,
                float* opac, float* shs, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_prepare_impl(ctx, n_assets, assets, M, means, scales, rots, opac, shs, (cudaStream_t)stream);
}

int lrt_prepare_backward(lrt_ctx* ctx, int n_assets, const lrt_asset* assets, int M, const float* dL_dmeans,
                         const float* dL_dscales, const float* dL_drots, const float* dL_dopac, const float* dL_dshs, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_prepare_backward_impl(ctx, n_assets, assets, M, dL_dmeans, dL_dscales, dL_drots, dL_dopac, dL_dshs, (cudaStream_t)stream);
}

int lrt_range_rays(lrt_ctx* ctx, int H, int W, const float* inc_table, float inc_lo, float inc_hi, float pixel_offset,
                   float angle_offset, const float* sensor2world, float* ray_d, float* centre, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_range_rays_impl(ctx, H, W, inc_table, inc_lo, inc_hi, pixel_offset, angle_offset, sensor2world, nullptr, ray_d, centre, (cudaStream_t)stream);
}

int lrt_range_points(lrt_ctx* ctx, int H, int W, const float* inc_table, float inc_lo, float inc_hi, float pixel_offset,
                     float angle_offset, const float* sensor2world, const float* range_map, float* points, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    if (!range_map) { ctx->set_error("lrt_range_points: range_map is null"); return LRT_ERR_INVALID; }
    return lrt_range_rays_impl(ctx, H, W, inc_table, inc_lo, inc_hi, pixel_offset, angle_offset, sensor2world, range_map, points, nullptr, (cudaStream_t)stream);
}

int lrt_chamfer_forward(lrt_ctx* ctx, int b, int n, const float* xyz1, int m, const float* xyz2,
                        float* dist1, int32_t* idx1, float* dist2, int32_t* idx2, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_chamfer_forward_impl(ctx, b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, (cudaStream_t)stream);
}

int lrt_chamfer_backward(lrt_ctx* ctx, int b, int n, const float* xyz1, int m, const float* xyz2,
                         const float* grad_dist1, const float* grad_dist2, const int32_t* idx1, const int32_t* idx2,
                         float* grad_xyz1, float* grad_xyz2, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_chamfer_backward_impl(ctx, b, n, xyz1, m, xyz2, grad_dist1, grad_dist2, idx1, idx2, grad_xyz1, grad_xyz2, (cudaStream_t)stream);
}

int lrt_adam_step(lrt_ctx* ctx, int n_tensors, const lrt_adam_tensor* tensors, double beta1, double beta2, double eps, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_adam_step_impl(ctx, n_tensors, tensors, beta1, beta2, eps, (cudaStream_t)stream);
}

int lrt_set_sh_parts(lrt_ctx* ctx, int n_parts, const lrt_sh_part* parts, int M, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    if (n_parts == 0) { ctx->sh_parts_n = 0; return LRT_OK; }
    if (n_parts < 0 || n_parts > LRT_MAX_ASSETS || !parts || M < 1 || M > 16) { ctx->set_error("lrt_set_sh_parts: need 0..LRT_MAX_ASSETS parts and 1 <= M <= 16"); return LRT_ERR_INVALID; }
    ShTab tab;
    long long total = 0;
    int vec = 1, gvec = 1;
    for (int k = 0; k < n_parts; k++) {
        const lrt_sh_part& p = parts[k];
        if (p.P <= 0 || !p.features_dc || (M > 1 && !p.features_rest)) { ctx->set_error("lrt_set_sh_parts: part with P <= 0 or a null tensor"); return LRT_ERR_INVALID; }
        tab.part[k].first = (int)total; tab.part[k].P = p.P; tab.part[k].dc = p.features_dc; tab.part[k].rest = p.features_rest;
        tab.part[k].d_dc = p.d_features_dc; tab.part[k].d_rest = p.d_features_rest;
        if (reinterpret_cast<uintptr_t>(p.features_rest) & 15) vec = 0;
        if (p.d_features_rest && (reinterpret_cast<uintptr_t>(p.d_features_rest) & 15)) gvec = 0;
        ctx->sh_parts[k] = p;
        total += p.P;
        if (total > 0x7fffffffLL / 64) { ctx->set_error("lrt_set_sh_parts: too many Gaussians"); return LRT_ERR_INVALID; }
    }
    tab.n = n_parts; tab.M = M;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LRT_ERR_CUDA;
    cudaError_t e = ctx->reserve(ctx->sh_tab, sizeof(ShTab));
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->sh_tab.p, &tab, sizeof(int) * 2 + sizeof(ShPartDev) * (size_t)n_parts, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { ctx->set_error("lrt_set_sh_parts", e); return LRT_ERR_CUDA; }
    ctx->sh_parts_n = n_parts; ctx->sh_parts_P = (int)total; ctx->sh_parts_M = M; ctx->sh_parts_vec = vec; ctx->sh_parts_grad_vec = gvec;
    return LRT_OK;
}

int lrt_compact_rows(lrt_ctx* ctx, int n_rows, const uint8_t* keep, int n_tensors, const lrt_row_tensor* tensors, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_compact_rows_impl(ctx, n_rows, keep, n_tensors, tensors, (cudaStream_t)stream);
}

int lrt_densify_rows(lrt_ctx* ctx, int P, const uint8_t* clone_mask, const uint8_t* split_mask, int n_clone, int n_split, int N,
                     const float* samples, const float* rotation, int n_tensors, const lrt_row_tensor* tensors, void* stream)
{
    if (!ctx) return LRT_ERR_INVALID;
    return lrt_densify_rows_impl(ctx, P, clone_mask, split_mask, n_clone, n_split, N, samples, rotation, n_tensors, tensors, (cudaStream_t)stream);
}

int lrt_set_option(lrt_ctx* ctx, int option, int value)
{
    if (!ctx) return LRT_ERR_INVALID;
    switch (option) {
    case LRT_OPT_FORWARD_KERNEL: if (value < 0 || value > 4) break; ctx->opt_forward_kernel = value; return LRT_OK;
    case LRT_OPT_RAY_GRID_WIDTH: if (value < 0) break; ctx->opt_ray_grid_w = value; return LRT_OK;
    case LRT_OPT_BEAM_CELL_PCT: if (value < 10 || value > 1000) break; ctx->opt_beam_cell_pct = value; return LRT_OK;
    case LRT_OPT_SORT_RAYS: if (value != 0 && value != 1) break; ctx->opt_sort_rays = value; return LRT_OK;
    case LRT_OPT_KERNEL_TIMING: if (value != 0 && value != 1) break; ctx->opt_kernel_timing = value; return LRT_OK;
    case LRT_OPT_WAVEFRONT_SHADE: if (value < 0 || value > 3) break; ctx->opt_wavefront_shade = value; return LRT_OK;
    case LRT_OPT_BACKWARD_KERNEL: if (value < 0 || value > 2) break; ctx->opt_backward_kernel = value; return LRT_OK;
    case LRT_OPT_BIN_CAP: if (value != 0 && (value < 512 || value > 16384 || (value & (value - 1)))) break; ctx->opt_bin_cap = value; return LRT_OK;
    case LRT_OPT_SORT_KEY_BITS: if (value != 16 && value != 24 && value != 32) break; ctx->opt_sort_key_bits = value; return LRT_OK;
    case LRT_OPT_MORTON_BITS: if (value != 30 && value != 32 && value != 63) break; ctx->opt_morton_bits = value; return LRT_OK;
    case LRT_OPT_VECTOR_ATOMICS: if (value != 0 && value != 1) break; ctx->opt_vector_atomics = value; return LRT_OK;
    case LRT_OPT_SPLIT_FUSED: if (value != 0 && value != 1) break; ctx->opt_split_fused = value; return LRT_OK;
    case LRT_OPT_TRIANGLE_DEPTH: if (value != 0 && value != 1) break; ctx->opt_triangle_depth = value; return LRT_OK;
    default: break;
    }
    ctx->set_error("lrt_set_option: unknown option or value out of range");
    return LRT_ERR_INVALID;
}

int lrt_get_info(const lrt_ctx* ctx, lrt_info* out)
{
    if (!ctx || !out) return LRT_ERR_INVALID;
    out->P = ctx->built ? ctx->P : 0;
    out->levels = ctx->levels;
    out->nodes = ctx->n_nodes;
    out->bytes_records = (int64_t)sizeof(SurfelRec) * ctx->P_pad;
    out->bytes_nodes = (int64_t)sizeof(Node8) * ctx->n_nodes;
    out->bytes_workspace = (int64_t)ctx->total_bytes();
    out->builds = ctx->builds; out->refits = ctx->refits;
    out->kernel_launches = ctx->launches;
    return LRT_OK;
}

/* Live per-kernel device times (LRT_OPT_KERNEL_TIMING = 1): sums the CUDA-event spans recorded since the last call,
 * per kernel name. Synchronises the events. names_out: `cap` slots of 32 chars; ms_out / count_out: `cap` entries.
 * Returns the number of distinct kernels (<= cap) or a negative status. */
int lrt_get_kernel_times(lrt_ctx* ctx, char* names_out, float* ms_out, int* count_out, int cap)
{
    if (!ctx || !names_out || !ms_out || !count_out || cap <= 0) return LRT_ERR_INVALID;
    std::map<std::string, std::pair<double, int>> acc;
    for (auto& t : ctx->spans) {
        float ms = 0.f;
        if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            auto& e = acc[t.name]; e.first += ms; e.second += 1;
        }
        ctx->event_pool.push_back(t.a); ctx->event_pool.push_back(t.b);
    }
    ctx->spans.clear();
    int n = 0;
    for (auto& kv : acc) {
        if (n >= cap) break;
        snprintf(names_out + 32 * n, 32, "%s", kv.first.c_str());
        ms_out[n] = (float)kv.second.first; count_out[n] = kv.second.second; n++;
    }
    return n;
}

/* development aid: the context's 16 work counters of the last forward (wavefront: [0..7] items per level,
 * [8] rays handed to the per-ray fallback). Synchronises the device. */
int lrt_debug_counters(const lrt_ctx* ctx, int* out)
{
    if (!ctx || !out) return LRT_ERR_INVALID;
    for (int i = 0; i < 16; i++) out[i] = 0;
    if (!ctx->counter.p || ctx->counter.cap < sizeof(int) * 16) return LRT_OK;
    if (cudaDeviceSynchronize() != cudaSuccess) return LRT_ERR_CUDA;
    return cudaMemcpy(out, ctx->counter.p, sizeof(int) * 16, cudaMemcpyDeviceToHost) == cudaSuccess ? LRT_OK : LRT_ERR_CUDA;
}

/* development statistics (non-zero only in the -DLRT_STATS build); out = 16 host uint64 */
int lrt_debug_stats(unsigned long long* out, int reset)
{
#ifdef LRT_STATS
    if (out && cudaMemcpyFromSymbol(out, g_lrt_stats, sizeof(unsigned long long) * 16) != cudaSuccess) return LRT_ERR_CUDA;
    if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(g_lrt_stats, z, sizeof(z)) != cudaSuccess) return LRT_ERR_CUDA; }
#else
    if (out) for (int i = 0; i < 16; i++) out[i] = 0;
    (void)reset;
#endif
    return LRT_OK;
}

int lrt_get_permutation(const lrt_ctx* ctx, int32_t* perm_out, void* stream)
{
    if (!ctx || !perm_out || !ctx->built) return LRT_ERR_STATE;
    cudaError_t e = cudaMemcpyAsync(perm_out, ctx->perm_a.p, sizeof(int32_t) * (size_t)ctx->P, cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream);
    return e == cudaSuccess ? LRT_OK : LRT_ERR_CUDA;
}

} // extern "C"
