// lrt_forward.cu — forward device program: k-buffer rounds + front-to-back compositing.
//
// Replaces __raygen__ot of the reference's forward pipeline
// (submodules/diff-lidar-tracer/optix_tracer/forward.cu:146-308) and TraceSurfelsCUDA
// (trace_surfels.cpp:151-265). One thread per ray; a round = trace_round() (lrt_trace.cuh);
// the compositing loop follows forward.cu:201-292 rule for rule:
//   depth = t' + base; skip depth < 0.2; skip cos == 0; alpha = min(.99, o exp(-(u^2+v^2)/2));
//   skip alpha < 1/255; stop BEFORE the hit that would take T below 1e-4; next round starts at
//   last depth + 1e-5 when the buffer was full.
// Additionally records, per ray, the ordered list of contributing surfels (id, depth) so the
// backward pass can replay it instead of traversing again.
#include "lrt_ctx.cuh"
#include "lrt_trace.cuh"

namespace {

__device__ __forceinline__ void load_sh(const float* __restrict__ shs, int g, int M, int nb, float* sh)
{
    const float* p = shs + (size_t)g * M * 3;
    if ((M & 3) == 0 && ((reinterpret_cast<uintptr_t>(shs) & 15) == 0)) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const int n4 = (nb * 3 + 3) >> 2;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            if (i < n4) { const float4 v = ld_f4(p4 + i); sh[4 * i] = v.x; sh[4 * i + 1] = v.y; sh[4 * i + 2] = v.z; sh[4 * i + 3] = v.w; }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 48; i++) if (i < nb * 3) sh[i] = ld_f(p + i);
    }
}

__global__ void __launch_bounds__(128)
k_forward(BvhView bvh, int R, const float* __restrict__ ray_o, int ray_o_stride, const float* __restrict__ ray_d,
          const float* __restrict__ bg, const float* __restrict__ shs, int D, int M,
          float* __restrict__ out, float* __restrict__ accum_w, int32_t* __restrict__ hit_gidx,
          float* __restrict__ hit_t, int32_t* __restrict__ hit_cnt, int cap, int32_t* __restrict__ slot_cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float o[3] = {ray_o[(size_t)r * ray_o_stride], ray_o[(size_t)r * ray_o_stride + 1], ray_o[(size_t)r * ray_o_stride + 2]};
    const float d[3] = {ray_d[3 * (size_t)r], ray_d[3 * (size_t)r + 1], ray_d[3 * (size_t)r + 2]};
    const float dl = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float dirn[3] = {d[0] / dl, d[1] / dl, d[2] / dl};
    const int nb = (D + 1) * (D + 1);

    float C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f, W = 0.f, T = 1.f, testT = 1.f, base = 0.f, dpt = 0.f;
    int ncontrib = 0, nslots = 0, last = -1;
#ifdef LRT_STATS
    int node_visits = 0;
#endif
    for (;;) {
        RaySetup rs;
        ray_setup(rs, o, d, base);
        unsigned long long kb[LRT_KBUF];
#ifdef LRT_STATS
        const int n = trace_round(bvh, rs, kb, node_visits);
#else
        const int n = trace_round(bvh, rs, kb);
#endif
        unsigned long long hits[LRT_KBUF];
#pragma unroll
        for (int i = 0; i < LRT_KBUF; i++) hits[i] = kb[i];
        bool terminated = false;
        for (int i = 0; i < n; i++) {
            const unsigned long long key = hits[i];
            const int g = (int)(unsigned)(key & 0xffffffffull);
            nslots++;
            dpt = __uint_as_float((unsigned)(key >> 32)) + base;                      // forward.cu:212
            if (dpt < LRT_MIN_T) continue;                                            // :214
            const float x0 = o[0] + dpt * d[0], x1 = o[1] + dpt * d[1], x2 = o[2] + dpt * d[2];
            // :220-224 — a re-based round can find the previous round's last surfel again at t' ~ +0
            // (the 1e-5 step is below one ulp of the depth beyond 128 m): it must not composite twice
            if (g == last) continue;
            last = g;
            const int prim = __ldg(bvh.iperm + g);
            const float4 a0 = ld_f4(&bvh.rec[prim].r0), a1 = ld_f4(&bvh.rec[prim].r1);
            const float4 a2 = ld_f4(&bvh.rec[prim].r2), a3 = ld_f4(&bvh.rec[prim].r3);
            const float r0 = x0 - a0.x, r1 = x1 - a0.y, r2 = x2 - a0.z;
            const float u = a1.x * r0 + a1.y * r1 + a1.z * r2;                        // :139
            const float v = a2.x * r0 + a2.y * r1 + a2.z * r2;
            const float cosv = -((a0.x - o[0]) * a3.x + (a0.y - o[1]) * a3.y + (a0.z - o[2]) * a3.z);
            if (cosv == 0.0f) continue;                                               // :233-237
            const float rho = u * u + v * v;
            const float power = -0.5f * rho;
            if (power > 0.0f) continue;
            const float G = expf(power);
            const float alpha = fminf(LRT_ALPHA_MAX, a1.w * G);                       // :249
            if (alpha < 1.0f / 255.0f) continue;
            testT = T * (1.0f - alpha);
            if (testT < LRT_T_MIN) { terminated = true; break; }                      // :253-257
            const float w = alpha * T;
            float sh[48], c[3]; bool cl;
            load_sh(shs, g, M, nb, sh);
            sh_colour<false>(D, dirn, sh, c, cl, nullptr);
            C0 += w * c[0]; C1 += w * c[1]; C2 += w * c[2];
            Dp += w * dpt; W += w;
            atomicAdd(accum_w + g, w);                                                // :272
            if (hit_gidx != nullptr && ncontrib < cap) {
                hit_gidx[(size_t)ncontrib * R + r] = g;
                hit_t[(size_t)ncontrib * R + r] = dpt;
            }
            ncontrib++;
            T = testT;
        }
        if (terminated || testT < LRT_T_MIN || n < LRT_KBUF) break;                   // :282-285
        base = (float)((double)dpt + LRT_STEP_EPS);                                   // :288
    }
    float* op = out + (size_t)LRT_NCH * r;
    op[0] = C0 + T * bg[0]; op[1] = C1 + T * bg[1]; op[2] = C2 + T * bg[2];          // :296-305
    op[3] = Dp; op[4] = W; op[5] = 0.f; op[6] = 0.f; op[7] = 0.f; op[8] = T;
    if (hit_cnt) hit_cnt[r] = ncontrib;
#ifdef LRT_STATS
    if (slot_cnt) slot_cnt[r] = (nslots & 0xffff) | (min(node_visits, 32767) << 16);   // development statistics build
#else
    if (slot_cnt) slot_cnt[r] = nslots;
#endif
}

} // namespace

int lrt_forward_impl(lrt_ctx* ctx, int R, const float* ray_o, int ray_o_stride, const float* ray_d,
                     const float* bg, int P, const float* means, const float* scales, const float* rots,
                     const float* opac, const float* shs, int D, int M, float mod,
                     float* out, float* accum_w, int32_t* hit_gidx, float* hit_t, int32_t* hit_cnt,
                     int cap, int32_t* slot_cnt, cudaStream_t s)
{
    (void)means; (void)scales; (void)rots; (void)opac;
    if (!ctx->built) { ctx->set_error("lrt_forward: no acceleration structure (call lrt_build first)"); return LRT_ERR_STATE; }
    if (P != ctx->P) { ctx->set_error("lrt_forward: P differs from the built structure"); return LRT_ERR_STATE; }
    if (mod != ctx->scale_modifier) { ctx->set_error("lrt_forward: scale_modifier differs from the built structure"); return LRT_ERR_STATE; }
    if (R < 0 || !ray_d || !ray_o || !bg || !shs || !out || !accum_w) { ctx->set_error("lrt_forward: null argument"); return LRT_ERR_INVALID; }
    if (ray_o_stride != 0 && ray_o_stride != 3) { ctx->set_error("lrt_forward: ray_o_stride must be 0 or 3"); return LRT_ERR_INVALID; }
    if (D < 0 || D > 3 || M < (D + 1) * (D + 1)) { ctx->set_error("lrt_forward: need 0 <= D <= 3 and M >= (D+1)^2"); return LRT_ERR_INVALID; }
    if ((hit_gidx == nullptr) != (hit_t == nullptr) || (hit_gidx && (cap <= 0 || !hit_cnt))) {
        ctx->set_error("lrt_forward: hit_gidx, hit_t and hit_cnt go together and need cap > 0"); return LRT_ERR_INVALID;
    }
    LRT_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LRT_CUDA_TRY(ctx, cudaMemsetAsync(accum_w, 0, sizeof(float) * (size_t)P, s));
    if (R == 0) return LRT_OK;
    const int TB = 128;
    k_forward<<<(R + TB - 1) / TB, TB, 0, s>>>(ctx->view(), R, ray_o, ray_o_stride, ray_d, bg, shs, D, M, out, accum_w,
                                               hit_gidx, hit_t, hit_cnt, cap, slot_cnt);
    ctx->launches += 1;
    LRT_CUDA_TRY(ctx, cudaGetLastError());
    return LRT_OK;
}
