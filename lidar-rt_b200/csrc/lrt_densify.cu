// lrt_densify.cu — densify / prune as row compaction over the packed parameter and optimiser-state tensors.
//
// The reference restructures a GaussianModel with torch indexing and concatenation, one tensor at a time, each step allocating
// every parameter and both Adam moments again (lib/scene/gaussian_model.py):
//   prune_points / _prune_optimizer (:235-270)            t = t[valid_mask] for 6 parameters, 12 moments, 3 statistics
//   densify_and_clone (:338-352)                          cat(t, t[clone_mask])            + cat(moment, zeros)
//   densify_and_split (:311-336)                          cat(t, children(t[split_mask]))  + cat(moment, zeros), then prune the parents
//   densify_and_prune (:354-407)                          the three above + a final prune by opacity / size
// Here the same result — the same surviving rows in the same order, moved bit for bit — comes from two primitives:
//   lrt_compact_rows   stable compaction of ANY number of row-major tensors by one keep-mask: one scan + one launch
//   lrt_densify_rows   clone + split + removal of the split parents in ONE pass: the output layout is the one the reference's
//                      three steps leave — [rows that are not split, in order | clones, in order | split children, N blocks
//                      in .repeat(N, 1) order] — with the children's position and scale computed in the kernel
//                      (build_rotation(rotation) . sample + xyz; log(exp(scaling) / (0.8 N))) and new moment rows zero.
// Random numbers stay with the caller: the normal samples of the split (torch.normal in the reference) are an input.
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include "lrt_ctx.cuh"

namespace {

struct RowTable { int n; int pad; lrt_row_tensor t[LRT_MAX_ROW_TENSORS]; };

struct MaskToInt {
    const unsigned char* m; int invert;
    __host__ __device__ int operator()(int i) const { return (m[i] != 0) != (invert != 0) ? 1 : 0; }
};

__global__ void __launch_bounds__(256) k_compact_rows(const __grid_constant__ RowTable tab, int n_rows, const unsigned char* __restrict__ keep,
                                                      const int* __restrict__ pos)
{
    const lrt_row_tensor t = tab.t[blockIdx.y];
    const int rf = t.row_floats;
    const unsigned n = (unsigned)n_rows * (unsigned)rf;             // the host checks n_rows * row_floats < 2^31: 32-bit index arithmetic
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const unsigned i = e / (unsigned)rf, c = e - i * (unsigned)rf;
        if (keep[i]) t.dst[(size_t)pos[i] * rf + c] = t.src[e];
    }
}

// build_rotation (lib/utils/general_utils.py:176-197) row `c` of R(q) dotted with s
__device__ __forceinline__ float rot_row_dot(const float* q_, int c, const float* s)
{
    const float n = sqrtf(q_[0] * q_[0] + q_[1] * q_[1] + q_[2] * q_[2] + q_[3] * q_[3]);
    const float r = q_[0] / n, x = q_[1] / n, y = q_[2] / n, z = q_[3] / n;
    float R0, R1, R2;
    if (c == 0) { R0 = 1.f - 2.f * (y * y + z * z); R1 = 2.f * (x * y - r * z); R2 = 2.f * (x * z + r * y); }
    else if (c == 1) { R0 = 2.f * (x * y + r * z); R1 = 1.f - 2.f * (x * x + z * z); R2 = 2.f * (y * z - r * x); }
    else { R0 = 2.f * (x * z - r * y); R1 = 2.f * (y * z + r * x); R2 = 1.f - 2.f * (x * x + y * y); }
    return (R0 * s[0] + R1 * s[1]) + R2 * s[2];
}

struct DensifyArgs {
    int P, n_keep, n_clone, n_split, N;
    const unsigned char* clone; const unsigned char* split;
    const int* pos_keep; const int* pos_clone; const int* pos_split;
    const float* samples;          // (N * n_split, 3): child b of the j-th split row at row b * n_split + j
    const float* rotation;         // (P, 4) raw quaternions of the parents
    float scale_div;               // 0.8 * N
};

__global__ void __launch_bounds__(256) k_densify_rows(const __grid_constant__ RowTable tab, const DensifyArgs a)
{
    const lrt_row_tensor t = tab.t[blockIdx.y];
    const int rf = t.row_floats;
    const unsigned n = (unsigned)a.P * (unsigned)rf;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const unsigned i = e / (unsigned)rf; const int c = (int)(e - i * (unsigned)rf);
        const float v = t.src[e];
        const bool sp = a.split[i] != 0;
        if (!sp) t.dst[(size_t)a.pos_keep[i] * rf + c] = v;                                  // survivors keep their order (and their moments)
        if (a.clone[i]) t.dst[(size_t)(a.n_keep + a.pos_clone[i]) * rf + c] = t.kind == LRT_ROW_ZERO_NEW ? 0.0f : v;
        if (sp) {
            const int j = a.pos_split[i];
            for (int b = 0; b < a.N; b++) {
                float w = v;
                if (t.kind == LRT_ROW_ZERO_NEW) w = 0.0f;
                else if (t.kind == LRT_ROW_XYZ) {                                            // gaussian_model.py:327
                    const float* s = a.samples + 3 * ((size_t)b * a.n_split + j);
                    const float q[4] = {a.rotation[4 * (size_t)i], a.rotation[4 * (size_t)i + 1], a.rotation[4 * (size_t)i + 2], a.rotation[4 * (size_t)i + 3]};
                    w = rot_row_dot(q, c, s) + v;
                } else if (t.kind == LRT_ROW_SCALING) {                                      // :328: log(exp(s) / (0.8 N))
                    w = logf(expf(v) / a.scale_div);
                }
                t.dst[(size_t)(a.n_keep + a.n_clone + (size_t)b * a.n_split + j) * rf + c] = w;
            }
        }
    }
}

int fill_rows(lrt_ctx* ctx, int n_rows, int n_tensors, const lrt_row_tensor* tensors, RowTable& tab, const char* who)
{
    if (n_tensors <= 0 || n_tensors > LRT_MAX_ROW_TENSORS || !tensors) { ctx->set_error((std::string(who) + ": need 1..LRT_MAX_ROW_TENSORS tensors").c_str()); return LRT_ERR_INVALID; }
    for (int k = 0; k < n_tensors; k++) {
        if (!tensors[k].src || !tensors[k].dst || tensors[k].row_floats <= 0) { ctx->set_error((std::string(who) + ": tensor with a null pointer or row_floats <= 0").c_str()); return LRT_ERR_INVALID; }
        if ((long long)n_rows * tensors[k].row_floats >= 0x7fffffffLL - 148 * 8 * 256) { ctx->set_error((std::string(who) + ": tensor too large (rows x row_floats must stay below 2^31)").c_str()); return LRT_ERR_INVALID; }
        tab.t[k] = tensors[k];
    }
    tab.n = n_tensors; tab.pad = 0;
    return LRT_OK;
}

cudaError_t scan_mask(lrt_ctx* ctx, const unsigned char* mask, int invert, int n, int* pos, cudaStream_t s)
{
    auto it = thrust::make_transform_iterator(thrust::make_counting_iterator(0), MaskToInt{mask, invert});     // the mask (or its negation) as 0 / 1
    size_t tb = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tb, it, pos, n, s);
    if (e != cudaSuccess) return e;
    e = ctx->reserve(ctx->dn_tmp, tb);
    if (e != cudaSuccess) return e;
    return cub::DeviceScan::ExclusiveSum(ctx->dn_tmp.p, tb, it, pos, n, s);
}

} // namespace

int lrt_compact_rows_impl(lrt_ctx* ctx, int n_rows, const unsigned char* keep, int n_tensors, const lrt_row_tensor* tensors, cudaStream_t s)
{
    if (n_rows < 0 || (n_rows > 0 && !keep)) { ctx->set_error("lrt_compact_rows: null mask"); return LRT_ERR_INVALID; }
    RowTable tab;
    const int rc = fill_rows(ctx, n_rows, n_tensors, tensors, tab, "lrt_compact_rows");
    if (rc != LRT_OK) return rc;
    if (n_rows == 0) return LRT_OK;
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->dn_pos, sizeof(int) * 3 * ((size_t)n_rows + 1)));
    int* pos = (int*)ctx->dn_pos.p;
    LRT_CUDA_TRY(ctx, scan_mask(ctx, keep, 0, n_rows, pos, s));
    ctx->span_begin("k_compact_rows", s);
    k_compact_rows<<<dim3(148 * 8, n_tensors), 256, 0, s>>>(tab, n_rows, keep, pos);
    ctx->span_end(s);
    ctx->launches += 2;
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}

int lrt_densify_rows_impl(lrt_ctx* ctx, int P, const unsigned char* clone_mask, const unsigned char* split_mask, int n_clone, int n_split,
                          int N, const float* samples, const float* rotation,
                          int n_tensors, const lrt_row_tensor* tensors, cudaStream_t s)
{
    if (P <= 0 || !clone_mask || !split_mask || n_clone < 0 || n_split < 0 || N < 1 || n_split > P || n_clone > P) { ctx->set_error("lrt_densify_rows: bad sizes or null masks"); return LRT_ERR_INVALID; }
    if (n_split > 0 && (!samples || !rotation)) { ctx->set_error("lrt_densify_rows: a split needs samples and rotation"); return LRT_ERR_INVALID; }
    RowTable tab;
    const int rc = fill_rows(ctx, P, n_tensors, tensors, tab, "lrt_densify_rows");
    if (rc != LRT_OK) return rc;
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LRT_CUDA_TRY(ctx, ctx->reserve(ctx->dn_pos, sizeof(int) * 3 * ((size_t)P + 1)));
    int* pos = (int*)ctx->dn_pos.p;
    DensifyArgs a;
    a.P = P; a.n_keep = P - n_split; a.n_clone = n_clone; a.n_split = n_split; a.N = N;
    a.clone = clone_mask; a.split = split_mask;
    a.pos_keep = pos; a.pos_clone = pos + (P + 1); a.pos_split = pos + 2 * ((size_t)P + 1);
    a.samples = samples; a.rotation = rotation;
    a.scale_div = 0.8f * (float)N;
    LRT_CUDA_TRY(ctx, scan_mask(ctx, split_mask, 1, P, pos, s));
    LRT_CUDA_TRY(ctx, scan_mask(ctx, clone_mask, 0, P, pos + (P + 1), s));
    LRT_CUDA_TRY(ctx, scan_mask(ctx, split_mask, 0, P, pos + 2 * ((size_t)P + 1), s));
    ctx->span_begin("k_densify_rows", s);
    k_densify_rows<<<dim3(148 * 8, n_tensors), 256, 0, s>>>(tab, a);
    ctx->span_end(s);
    ctx->launches += 4;
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}
