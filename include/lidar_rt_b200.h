/*
 * lidar_rt_b200.h — C ABI of the B200-native differentiable LiDAR Gaussian ray tracer.
 *
 * This is the drop-in boundary for the ONE hot path of zju3dv/LiDAR-RT: everything that sits
 * behind the reference's pybind11 module `diff_lidar_tracer._C`
 * (submodules/diff-lidar-tracer/ext.cpp:17-22, trace_surfels.h:21-77). Plain pointers, sizes and
 * a cudaStream_t (as void*); no torch types. All pointers are DEVICE pointers on the context's
 * device unless stated otherwise; all floats are fp32; tensors are dense row-major.
 *
 * Error model: every function returns 0 on success or a negative lrt_status; the message is
 * available from lrt_last_error(). (The reference only prints CUDA/OptiX failures —
 * optix_tracer/common.h:38-50 — here they are returned.)
 *
 * Threading: a context is thread-compatible (one call at a time per context). Calls enqueue work
 * on `stream` and do not synchronise the device (the reference ends every call with
 * cudaStreamSynchronize, trace_surfels.cpp:260,382). Workspace growth may call cudaMalloc.
 */
#ifndef LIDAR_RT_B200_H
#define LIDAR_RT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRT_VERSION 100
#define LRT_NUM_CHANNELS 9        /* optix_tracer/config.h:19-24: rgb(3) depth accum normal(3) finalT */
#define LRT_CHUNK 16              /* optix_tracer/config.h:16 CHUNK_SIZE */

typedef struct lrt_ctx lrt_ctx;

enum lrt_status {
    LRT_OK = 0,
    LRT_ERR_INVALID = -1,         /* bad argument (the reference raises via AT_ERROR, trace_surfels.cpp:53-58) */
    LRT_ERR_CUDA = -2,            /* a CUDA runtime call failed */
    LRT_ERR_STATE = -3            /* no acceleration structure / size mismatch with the built one */
};

enum lrt_flags {
    LRT_FLAG_FIX_BG_GRAD = 1      /* drop the duplicated background term of backward.cu:595-598 */
};

/* Replaces _C.OptiXStateWrapper(pkg_dir) (optix_wrapper.cpp:177-233): creates the per-device
 * context that owns the acceleration structure and workspace. */
int lrt_ctx_create(int device, lrt_ctx** out_ctx);
int lrt_ctx_destroy(lrt_ctx* ctx);
/* Message of the last failing call on this context (ctx == NULL: last lrt_ctx_create failure). */
const char* lrt_last_error(const lrt_ctx* ctx);
int lrt_version(void);

/* Replaces build2DRectangle (lib/utils/primitive_utils.py:182-224) + _C.build_acceleration_structure
 * with rebuild=1 (trace_surfels.cpp:46-148): derives the proxy quad of every Gaussian straight from
 * its parameters and builds the LBVH (Morton sort + 8-wide implicit hierarchy) over the quad AABBs.
 *   means (P,3)  scales (P,2) >0  rots (P,4) w-first, need not be unit  opac (P) in (0,1)
 * The arrays are only read during the call. */
int lrt_build(lrt_ctx* ctx, int P, const float* means, const float* scales, const float* rots,
              const float* opac, float scale_modifier, void* stream);

/* The rebuild=0 / OPTIX_BUILD_OPERATION_UPDATE path (trace_surfels.cpp:70-73): keeps the Morton
 * order of the last lrt_build and recomputes records and boxes bottom-up. Same P required. */
int lrt_refit(lrt_ctx* ctx, int P, const float* means, const float* scales, const float* rots,
              const float* opac, float scale_modifier, void* stream);

/* Replaces _C.trace_surfels (trace_surfels.cpp:151-265; device program forward.cu:146-356).
 *   R rays; ray_o_stride = 3 (origins (R,3)) or 0 (one shared origin — the expanded stride-0 view
 *   LiDARSensor.get_range_rays returns, lidar_sensor.py:400); ray_d (R,3); bg (3) device.
 *   Gaussian arrays as given to lrt_build/lrt_refit (same values); shs (P,M,3); D = active degree.
 *   shs may be NULL after lrt_set_sh_parts (rows read in place from the model's features_dc / features_rest).
 * Outputs (written entirely by the call):
 *   out (R,9)           channel map of config.h:19-24
 *   accum_w (P)         sum of blending weights per Gaussian (forward.cu:272)
 *   hit_gidx, hit_t     optional, (cap,R) each: contributing Gaussian ids (caller's indexing) and
 *                       depths in compositing order, slot k of ray r at [k*R + r]
 *   hit_aux             optional (needs the lists), (cap,R,4) 16-byte aligned: per recorded hit (alpha, c0, c1, c2) — the
 *                       blending opacity and SH colour the forward composited (c0 = -0.0 where channel 0 was clamped).
 *                       With it the backward needs neither a serial re-evaluation of the ray nor the SH coefficients.
 *   hit_cnt (R)         optional: number of contributing hits (may exceed cap; list is truncated)
 *   slot_cnt (R)        optional: k-buffer slots consumed (evaluated proxy hits)
 */
int lrt_forward(lrt_ctx* ctx, int R, const float* ray_o, int ray_o_stride, const float* ray_d,
                const float* bg, int P, const float* means, const float* scales, const float* rots,
                const float* opac, const float* shs, int D, int M, float scale_modifier,
                float* out, float* accum_w, int32_t* hit_gidx, float* hit_t, float* hit_aux, int32_t* hit_cnt,
                int cap, int32_t* slot_cnt, void* stream);

/* Replaces _C.trace_surfels_backward (trace_surfels.cpp:268-386; backward.cu:434-691).
 * fwd_out / dL_dout (R,9). If hit lists from the forward are given, rays with hit_cnt <= cap are
 * replayed from the list (no traversal; with hit_aux as two passes: a per-ray prefix pass over the recorded
 * (alpha, colour) and a pass with one thread per hit that scatters the gradients); the others — or all rays when the lists are NULL — are
 * re-traced through the current acceleration structure like the reference does.
 * Gradients (written entirely by the call, accumulated with float atomics):
 *   dL_dmeans (P,3)  dL_dshs (P,M,3)  dL_dopac (P)  dL_dscales (P,2)  dL_drots (P,4) */
int lrt_backward(lrt_ctx* ctx, int R, const float* ray_o, int ray_o_stride, const float* ray_d,
                 const float* bg, int P, const float* means, const float* scales, const float* rots,
                 const float* opac, const float* shs, int D, int M, float scale_modifier,
                 const float* fwd_out, const float* dL_dout,
                 const int32_t* hit_gidx, const float* hit_t, const float* hit_aux, const int32_t* hit_cnt, int cap,
                 float* dL_dmeans, float* dL_dshs, float* dL_dopac, float* dL_dscales,
                 float* dL_drots, int flags, void* stream);

/* ---- the step above the tracer: parameter activation + world transform + concatenation, fused ----
 * Replaces, in lib/gaussian_renderer/__init__.py:76-134, the per-asset accessor calls of the reference's GaussianModel
 * (lib/scene/gaussian_model.py:112-148: exp / sigmoid / normalize, xyz @ R^T + T, cat(features_dc, features_rest)),
 * the quaternion composition of dynamic scenes (general_utils.py:156-174) and the torch.cat over assets.
 * One lrt_asset per GaussianModel, in concatenation order (background first). All pointers are device pointers. */
#define LRT_MAX_ASSETS 128
typedef struct lrt_asset {
    int32_t P;                    /* Gaussians of this asset */
    int32_t compose_rotation;     /* 0: rotations = normalize(rotation)                       (:117-118, static scene / background)
                                     1: rotations = pose_quat (x) normalize(normalize(rotation))  (:119-130, actors of a dynamic scene) */
    const float* xyz;             /* (P,3)   _xyz            local frame */
    const float* scaling;         /* (P,2)   _scaling        log-scale, 8-byte aligned */
    const float* rotation;        /* (P,4)   _rotation       raw quaternion (w first), 16-byte aligned */
    const float* opacity;         /* (P,1)   _opacity        logit */
    const float* features_dc;     /* (P,1,3) _features_dc */
    const float* features_rest;   /* (P,M-1,3) _features_rest */
    const float* pose_T;          /* (3) translation of BoundingBox.frame[t], or NULL: world xyz = xyz (gaussian_model.py:133-138) */
    const float* pose_quat;       /* (4) quaternion of BoundingBox.frame[t] (R = build_rotation(q)), or NULL */
    /* lrt_prepare_backward only: where the leaf gradients go (each may be NULL = not wanted); written, not accumulated */
    float* d_xyz; float* d_scaling; float* d_rotation; float* d_opacity; float* d_features_dc; float* d_features_rest;
} lrt_asset;

/* assets: HOST array of n_assets descriptors. Outputs (device, Ptot = sum of P): means (Ptot,3), scales (Ptot,2) 8-byte aligned,
 * rots (Ptot,4) 16-byte aligned, opac (Ptot), shs (Ptot,M,3) — exactly what lrt_build / lrt_forward take. */
int lrt_prepare(lrt_ctx* ctx, int n_assets, const lrt_asset* assets, int M, float* means, float* scales, float* rots,
                float* opac, float* shs, void* stream);
/* VJP of lrt_prepare: gradients w.r.t. its five outputs in, leaf gradients out through the d_* pointers of the assets.
 * No gradient flows to the poses (plain tensors in the reference). */
int lrt_prepare_backward(lrt_ctx* ctx, int n_assets, const lrt_asset* assets, int M, const float* dL_dmeans,
                         const float* dL_dscales, const float* dL_drots, const float* dL_dopac, const float* dL_dshs, void* stream);

/* ---- ray generation / back-projection of a LiDAR range image ----
 * lrt_range_rays replaces LiDARSensor.get_range_rays (lib/scene/lidar_sensor.py:395-434): unit ray directions (H,W,3) in the
 * world frame and the shared origin (3) — pass it to lrt_forward with ray_o_stride = 0.
 * lrt_range_points replaces LiDARSensor.range2point (:325-393): world points (H,W,3) of a range map (H,W).
 *   inc_table: device (H) beam inclinations in ascending order (the reference's list form), or NULL to use the two bounds
 *   [inc_lo, inc_hi] (its 2-element form); sensor2world: device (4,4) row-major. */
int lrt_range_rays(lrt_ctx* ctx, int H, int W, const float* inc_table, float inc_lo, float inc_hi, float pixel_offset,
                   float angle_offset, const float* sensor2world, float* ray_d, float* centre, void* stream);
int lrt_range_points(lrt_ctx* ctx, int H, int W, const float* inc_table, float inc_lo, float inc_hi, float pixel_offset,
                     float angle_offset, const float* sensor2world, const float* range_map, float* points, void* stream);

/* ---- Chamfer distance between two point clouds ----
 * lrt_chamfer_forward replaces chamfer_3D.forward (lib/utils/chamfer3D/chamfer_cuda.cpp:17-19 -> chamfer3D.cu:135-155, kernel :11-133):
 *   xyz1 (b,n,3), xyz2 (b,m,3); for every point of xyz1 the squared distance to its nearest point of xyz2 and that point's index
 *   (the lowest index among exact ties, as the reference's strict comparisons give), and the same for xyz2 against xyz1.
 *   dist1 (b,n), idx1 (b,n) int32, dist2 (b,m), idx2 (b,m); written entirely by the call. An empty opposite cloud leaves zeros,
 *   like the reference's zero-initialised outputs. Inputs must be finite.
 * lrt_chamfer_backward replaces chamfer_3D.backward (chamfer_cuda.cpp:22-26 -> chamfer3D.cu:157-196): grad_xyz1 (b,n,3) and
 *   grad_xyz2 (b,m,3) from grad_dist1 (b,n), grad_dist2 (b,m) and the indices of the forward; written entirely by the call
 *   (the reference accumulates into zero-filled tensors), neighbour terms accumulated with float atomics. */
int lrt_chamfer_forward(lrt_ctx* ctx, int b, int n, const float* xyz1, int m, const float* xyz2,
                        float* dist1, int32_t* idx1, float* dist2, int32_t* idx2, void* stream);
int lrt_chamfer_backward(lrt_ctx* ctx, int b, int n, const float* xyz1, int m, const float* xyz2,
                         const float* grad_dist1, const float* grad_dist2, const int32_t* idx1, const int32_t* idx2,
                         float* grad_xyz1, float* grad_xyz2, void* stream);

/* ---- the optimiser step under the tracer's gradients ----
 * lrt_adam_step replaces the per-asset, per-group torch.optim.Adam(l, lr=0.0, eps=1e-15).step() calls of the reference
 * (lib/scene/gaussian_model.py:186-201 sets the groups up; train.py steps every asset's optimizer each iteration): every
 * parameter tensor of every asset is one row of `tensors` (HOST array) and all rows are updated by one launch, with the arithmetic
 * of torch's single-tensor Adam (amsgrad off, no weight decay):
 *   m = m + (g - m)(1 - beta1);  v = v beta2 + ((1 - beta2) g) g;  p = p + (-(lr / (1 - beta1^step)) m) / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 * beta1 / beta2 / eps are doubles because torch forms 1 - beta in double before rounding it to the tensor's type.
 * param / exp_avg / exp_avg_sq are updated in place; `step` is the 1-based count of THIS tensor's update (state re-created by
 * densification restarts at 1). */
typedef struct lrt_adam_tensor {
    float* param; const float* grad; float* exp_avg; float* exp_avg_sq;   /* device, n floats each */
    int64_t n;
    float lr;
    int32_t step;
} lrt_adam_tensor;
int lrt_adam_step(lrt_ctx* ctx, int n_tensors, const lrt_adam_tensor* tensors, double beta1, double beta2, double eps, void* stream);

/* ---- SH coefficients read, and differentiated, IN PLACE ----
 * The reference concatenates cat(features_dc, features_rest) of every asset into one (P, M, 3) tensor per render call
 * (gaussian_model.py:141-144, gaussian_renderer/__init__.py:105,131) and autograd splits the gradient back: 2 x 0.92 GB of copies per
 * training step at 2.4 M Gaussians. After lrt_set_sh_parts, lrt_forward may be called with shs == NULL and lrt_backward with
 * shs == NULL and dL_dshs == NULL: rows are then fetched from the parts (Gaussian g of the concatenation = row g - first_k of part k),
 * and SH gradients are accumulated straight into d_features_dc / d_features_rest (zero-filled by lrt_backward; NULL = not wanted).
 * parts: HOST array, copied by the call; the tensors must stay valid until the calls that use them have run. M as in lrt_forward.
 * n_parts == 0 unbinds. 16-byte aligned features_rest / d_features_rest bases enable the vector paths. */
typedef struct lrt_sh_part {
    int32_t P; int32_t reserved;
    const float* features_dc;     /* (P,1,3) */
    const float* features_rest;   /* (P,M-1,3) */
    float* d_features_dc;         /* lrt_backward: (P,1,3) or NULL */
    float* d_features_rest;       /* lrt_backward: (P,M-1,3) or NULL */
} lrt_sh_part;
int lrt_set_sh_parts(lrt_ctx* ctx, int n_parts, const lrt_sh_part* parts, int M, void* stream);

/* ---- densify / prune as row compaction (second half) ----
 * The reference restructures a GaussianModel one tensor at a time with torch indexing / torch.cat, re-allocating every parameter and
 * both Adam moments at each of prune_points, densify_and_clone, densify_and_split (lib/scene/gaussian_model.py:235-352). Here every
 * tensor that has one row per Gaussian — the six parameters, their exp_avg / exp_avg_sq, the densification statistics — is one
 * lrt_row_tensor of a HOST table and ONE call moves them all. dst tensors are allocated by the caller (row counts below). */
#define LRT_MAX_ROW_TENSORS 32
enum lrt_row_kind {
    LRT_ROW_COPY = 0,             /* new rows copy their parent's row (rotation, features, opacity) */
    LRT_ROW_ZERO_NEW = 1,         /* new rows are zero (optimiser moments: cat(state, zeros_like(extension)), :281-282) */
    LRT_ROW_XYZ = 2,              /* split children: build_rotation(rotation) . sample + xyz (:326-327); clones copy */
    LRT_ROW_SCALING = 3           /* split children: log(exp(scaling) / (0.8 N)) (:328); clones copy */
};
typedef struct lrt_row_tensor {
    const float* src;             /* (n_rows, row_floats) */
    float* dst;                   /* (rows out, row_floats) */
    int32_t row_floats;
    int32_t kind;                 /* lrt_row_kind; ignored by lrt_compact_rows */
} lrt_row_tensor;
/* prune_points (:253-270): dst = src[keep] for every tensor, order kept. keep: device (n_rows) bytes, non-zero = keep. */
int lrt_compact_rows(lrt_ctx* ctx, int n_rows, const uint8_t* keep, int n_tensors, const lrt_row_tensor* tensors, void* stream);
/* densify_and_clone + densify_and_split including the removal of the split parents (:311-352), in one pass. Output rows
 *   [ rows with split_mask == 0, in order | rows with clone_mask != 0, in order | N blocks of the split rows' children ]
 * = P - n_split + n_clone + N n_split rows (n_clone / n_split = the masks' population counts, which the caller has: the reference
 * calls .sum().item() on them too). samples: (N n_split, 3) normal samples in the reference's .repeat(N, 1) order (child b of the
 * j-th split row at row b n_split + j; the caller draws them: torch.normal(0, cat(exp(scaling), 0))); rotation: (P, 4) raw. */
int lrt_densify_rows(lrt_ctx* ctx, int P, const uint8_t* clone_mask, const uint8_t* split_mask, int n_clone, int n_split, int N,
                     const float* samples, const float* rotation, int n_tensors, const lrt_row_tensor* tensors, void* stream);

/* Tuning knobs; none of them changes results — except LRT_OPT_TRIANGLE_DEPTH, which selects the reference's literal proxy geometry.
 *   LRT_OPT_FORWARD_KERNEL  0 = one thread per ray, 1 = persistent threads with per-lane refill, 2 = 8 lanes per ray,
 *                           3 = breadth-first wavefront through the hierarchy + per-ray sort + compositing,
 *                           4 = shared-origin beam grid (default): when ray_o_stride == 0 the frame's rays are binned by
 *                               direction and one pass over the surfel records fills the per-ray candidate bins; frames
 *                               with per-ray origins take 3
 *   LRT_OPT_RAY_GRID_WIDTH  W > 0: the R rays of the next calls are a row-major (R / W, W) range image
 *                           (the (H, W, 3) tensors of the reference API); lets a warp take a 4 x 8 tile of
 *                           neighbouring rays. 0 = no structure known (default).
 *   LRT_OPT_VECTOR_ATOMICS  backward: 128-bit vector reductions where alignment allows (default 1)
 *   LRT_OPT_BACKWARD_KERNEL 0 = one thread per ray replays its hit list, 1 = one warp per ray, one hit per lane,
 *                           2 = per-ray prefix pass + one thread per hit (default; used when hit_aux is given, else 0)
 *   LRT_OPT_SORT_RAYS       1 = compositing and the backward replay take rays in order of descending list length, so the
 *                           32 lanes of a warp run loops of equal length (default 1)
 *   LRT_OPT_BEAM_CELL_PCT   beam grid: cell edge in percent of the size that gives one ray per cell (default 100)
 *   LRT_OPT_KERNEL_TIMING   1 = record CUDA events around every kernel launch (read with lrt_get_kernel_times)
 *   LRT_OPT_WAVEFRONT_SHADE wavefront compositing: 0 = one warp per ray, 1 = warp sort + one thread per ray,
 *                           2 = 1 with pipelined record loads and slot opacities computed on acceptance,
 *                           3 = split passes (default): sort + gather into a sorted record stream, then slots / colour / fold
 *   LRT_OPT_SPLIT_FUSED     split passes: 1 = the bin sort and the rounds' slot logic run in one kernel, one warp per ray, straight
 *                           from the surfel records (default); 0 = two kernels with a sorted record stream in between
 *   LRT_OPT_TRIANGLE_DEPTH  1 = the hits of a ray and their depths come from the reference's literal proxy, the two triangles
 *                           (v0,v1,v2), (v2,v3,v1) over the corners build2DRectangle rounds to fp32 (primitive_utils.py:203-221),
 *                           intersected in fp64 like the oracle's ORC_TRIANGLES mode, instead of the analytic quad |u|,|v| <= f:
 *                           what forward.cu:319 reads with optixGetRayTmax. Default 0. Applies to the default forward path
 *                           (kernel 4 / 3 with split passes); rays handed to the fallback paths keep the analytic quad.
 *   LRT_OPT_MORTON_BITS     32 = 32-bit keys, bits dealt to the axes so cells stay cubic (default); 63 = 21 bits/axis on
 *                           cubic cells; 30 = 10 bits/axis on the per-axis extent (takes effect at the next lrt_build)
 *   LRT_OPT_SORT_KEY_BITS   how many of the top bits of a 32-bit key the build's radix sort orders (16 = default, 24, 32): one 8-bit
 *                           pass each; surfels that agree on those bits keep the caller's order, results do not depend on it
 *   LRT_OPT_BIN_CAP         candidate-bin capacity per ray: 0 = automatic (default), or a power of two in [512, 16384]. Candidates
 *                           beyond it travel through the overflow list of the split passes; results do not depend on it */
enum lrt_option { LRT_OPT_FORWARD_KERNEL = 1, LRT_OPT_RAY_GRID_WIDTH = 2, LRT_OPT_VECTOR_ATOMICS = 3, LRT_OPT_MORTON_BITS = 4,
                  LRT_OPT_BACKWARD_KERNEL = 5, LRT_OPT_WAVEFRONT_SHADE = 6, LRT_OPT_KERNEL_TIMING = 7,
                  LRT_OPT_SORT_RAYS = 8, LRT_OPT_BEAM_CELL_PCT = 9, LRT_OPT_TRIANGLE_DEPTH = 10, LRT_OPT_SPLIT_FUSED = 11, LRT_OPT_SORT_KEY_BITS = 12, LRT_OPT_BIN_CAP = 13 };
int lrt_set_option(lrt_ctx* ctx, int option, int value);

/* Introspection for tests / benchmarks (host pointers). */
typedef struct lrt_info {
    int32_t P;                    /* Gaussians in the built structure */
    int32_t levels;               /* hierarchy levels (8-wide) */
    int64_t nodes;                /* total nodes */
    int64_t bytes_records;        /* HBM bytes of the sorted surfel records */
    int64_t bytes_nodes;          /* HBM bytes of the hierarchy */
    int64_t bytes_workspace;      /* total device bytes owned by the context */
    int64_t builds, refits;       /* counters */
    int32_t kernel_launches;      /* kernels launched by this library since context creation */
} lrt_info;
int lrt_get_info(const lrt_ctx* ctx, lrt_info* out);
/* Live per-kernel device times since the previous call (needs LRT_OPT_KERNEL_TIMING = 1); host pointers:
 * names_out = cap x 32 chars, ms_out / count_out = cap entries. Returns the number of kernels reported. */
int lrt_get_kernel_times(lrt_ctx* ctx, char* names_out, float* ms_out, int* count_out, int cap);
/* development aid: work counters of the last forward (wavefront: [0..7] items per level, [8] fallback rays) */
int lrt_debug_counters(const lrt_ctx* ctx, int* out);
/* development counters (all zero unless the library was built with -DLRT_STATS); out = 16 host uint64 */
int lrt_debug_stats(unsigned long long* out, int reset);
/* sorted position -> caller's Gaussian index, (P) int32 device copy */
int lrt_get_permutation(const lrt_ctx* ctx, int32_t* perm_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIDAR_RT_B200_H */
