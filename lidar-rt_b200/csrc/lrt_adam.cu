// lrt_adam.cu — one Adam step over every parameter tensor of every Gaussian asset, in one launch.
//
// SURVEY.md §8(f) N4. The reference gives each GaussianModel its own torch.optim.Adam with six parameter groups
// (lib/scene/gaussian_model.py:186-201: xyz, f_dc, f_rest, opacity, scaling, rotation; eps = 1e-15) and steps them one after
// the other (train.py: `gaussians.optimizer.step()` per asset): 41 assets x 6 tensors x ~8 elementwise kernels per iteration on a
// Waymo-dynamic scene, most of them over 10 000-Gaussian actors, i.e. launch-bound. Here the tensors are rows of one table
// and a block finds its (tensor, chunk) by binary search over the table's block prefix.
//
// Arithmetic: torch.optim.Adam's single-tensor path (torch/optim/adam.py, _single_tensor_adam; amsgrad = False,
// weight_decay = 0, maximize = False), operation for operation in fp32:
//   m    = m + (g - m) * (1 - beta1)                     exp_avg.lerp_(grad, 1 - beta1)
//   v    = v * beta2 + ((1 - beta2) * g) * g             exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
//   den  = sqrt(v) / sqrt(1 - beta2^t) + eps             (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
//   p    = p + (-(lr / (1 - beta1^t)) * m) / den         param.addcdiv_(exp_avg, denom, value = -step_size)
// The bias corrections are evaluated on the host in double, like torch does, and passed per tensor (each tensor has its
// own step count: densification re-creates state).
#include "lrt_ctx.cuh"

namespace {

constexpr int AD_TB = 256;
constexpr int AD_PER_THREAD = 8;
constexpr int AD_CHUNK = AD_TB * AD_PER_THREAD;
constexpr int AD_MAX_TENSORS = 512;           // per launch (the table travels as a kernel parameter: 512 x 56 B = 28 KB)

struct AdamRow {
    float* p; const float* g; float* m; float* v;
    long long n;
    float neg_step_size, bc2_sqrt;            // -(lr / (1 - beta1^t)), sqrt(1 - beta2^t)
    int block0;                               // first block of this tensor
    int pad;
};
struct AdamTable { int n_rows, n_blocks; float w1, beta2, w2, eps; AdamRow r[AD_MAX_TENSORS]; };

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamTable& t, const AdamRow& r)
{
    m = m + (g - m) * t.w1;
    v = v * t.beta2 + (t.w2 * g) * g;
    const float den = sqrtf(v) / r.bc2_sqrt + t.eps;
    p = p + (r.neg_step_size * m) / den;
}

__global__ void __launch_bounds__(AD_TB) k_adam(const __grid_constant__ AdamTable t)
{
    int lo = 0, hi = t.n_rows - 1;            // last row whose first block is <= blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t.r[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const AdamRow& r = t.r[lo];
    const long long base = (long long)(blockIdx.x - r.block0) * AD_CHUNK;
    const bool vec = ((reinterpret_cast<uintptr_t>(r.p) | reinterpret_cast<uintptr_t>(r.g) | reinterpret_cast<uintptr_t>(r.m) |
                       reinterpret_cast<uintptr_t>(r.v)) & 15) == 0;
#pragma unroll
    for (int k = 0; k < AD_PER_THREAD / 4; k++) {
        const long long i = base + 4 * ((long long)k * AD_TB + threadIdx.x);
        if (i >= r.n) continue;
        if (vec && i + 4 <= r.n) {
            float4 p = *reinterpret_cast<float4*>(r.p + i), m = *reinterpret_cast<float4*>(r.m + i), v = *reinterpret_cast<float4*>(r.v + i);
            const float4 g = __ldg(reinterpret_cast<const float4*>(r.g + i));
            adam_one(p.x, g.x, m.x, v.x, t, r); adam_one(p.y, g.y, m.y, v.y, t, r);
            adam_one(p.z, g.z, m.z, v.z, t, r); adam_one(p.w, g.w, m.w, v.w, t, r);
            *reinterpret_cast<float4*>(r.p + i) = p; *reinterpret_cast<float4*>(r.m + i) = m; *reinterpret_cast<float4*>(r.v + i) = v;
        } else {
            for (long long j = i; j < i + 4 && j < r.n; j++) {
                float p = r.p[j], m = r.m[j], v = r.v[j];
                adam_one(p, r.g[j], m, v, t, r);
                r.p[j] = p; r.m[j] = m; r.v[j] = v;
            }
        }
    }
}

} // namespace

int lrt_adam_step_impl(lrt_ctx* ctx, int n_tensors, const lrt_adam_tensor* tensors, double beta1, double beta2, double eps, cudaStream_t s)
{
    if (n_tensors < 0 || (n_tensors > 0 && !tensors)) { ctx->set_error("lrt_adam_step: bad table"); return LRT_ERR_INVALID; }
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(AdamTable) <= 32000, "the table must fit the kernel parameter space");
    AdamTable t;
    // torch converts the python scalars 1 - beta1, beta2, 1 - beta2, eps to the tensor dtype when the op runs
    t.w1 = (float)(1.0 - beta1); t.beta2 = (float)beta2; t.w2 = (float)(1.0 - beta2); t.eps = (float)eps;
    int k = 0;
    while (k < n_tensors) {
        t.n_rows = 0; t.n_blocks = 0;
        for (; k < n_tensors && t.n_rows < AD_MAX_TENSORS; k++) {
            const lrt_adam_tensor& e = tensors[k];
            if (e.n < 0 || e.step < 1 || (e.n > 0 && (!e.param || !e.grad || !e.exp_avg || !e.exp_avg_sq))) {
                ctx->set_error("lrt_adam_step: tensor with null pointer, negative size or step < 1"); return LRT_ERR_INVALID;
            }
            if (e.n == 0) continue;
            const long long nb = (e.n + AD_CHUNK - 1) / AD_CHUNK;
            if (t.n_blocks + nb > 0x7fffffffLL) break;
            AdamRow& r = t.r[t.n_rows++];
            r.p = e.param; r.g = e.grad; r.m = e.exp_avg; r.v = e.exp_avg_sq; r.n = e.n;
            const double bc1 = 1.0 - pow(beta1, (double)e.step), bc2 = 1.0 - pow(beta2, (double)e.step);
            r.neg_step_size = (float)(-((double)e.lr / bc1));
            r.bc2_sqrt = (float)sqrt(bc2);
            r.block0 = t.n_blocks; r.pad = 0;
            t.n_blocks += (int)nb;
        }
        if (t.n_rows == 0) { if (k < n_tensors && tensors[k].n > 0) { ctx->set_error("lrt_adam_step: tensor too large"); return LRT_ERR_INVALID; } continue; }
        ctx->span_begin("k_adam", s);
        k_adam<<<t.n_blocks, AD_TB, 0, s>>>(t);
        ctx->span_end(s);
        ctx->launches += 1;
        LRT_CUDA_TRY(ctx, cudaGetLastError());
    }
    return LRT_OK;
}
