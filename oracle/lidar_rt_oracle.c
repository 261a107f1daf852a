/*
 * oracle/lidar_rt_oracle.c — TEST INFRASTRUCTURE. NOT PART OF THE PRODUCT.
 *
 * A plain-C CPU restatement of the reference LiDAR-RT tracer hot path, used ONLY as the parity
 * checker (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).
 * Nothing under lidar-rt_b200/ may import, link or execute it.
 *
 * Parity pin: this file is validated against oracle/_ref (the reference's own forward.cu /
 * backward.cu compiled unmodified as host code over an OptiX stand-in, see oracle/build_ref.sh)
 * by tests/test_oracle_vs_ref.py and by the committed fixtures in tests/golden/ that
 * oracle/make_golden.py generated from oracle/_ref.
 *
 * What is restated (paths relative to /root/reference, DLT = submodules/diff-lidar-tracer):
 *   - proxy quad per Gaussian ........ lib/utils/primitive_utils.py:182-224 (build2DRectangle),
 *                                      lib/utils/general_utils.py:176-197 (build_rotation)
 *   - k-buffer any-hit + round loop .. DLT/optix_tracer/forward.cu:146-356, config.h:16-17
 *   - surfel response / compositing .. forward.cu:116-141, :201-305
 *   - SH colour ...................... forward.cu:67-111, auxiliary.h:23-40
 *   - backward (re-trace + VJPs) ..... backward.cu:434-691, :339-431, :123-291,
 *                                      auxiliary.h:389-433 (quat_to_rotmat_vjp)
 *
 * Arithmetic contract ("canonical form"): every expression is evaluated in `real`
 * (float unless -DORC_DOUBLE) as written, left to right, with NO fused contraction
 * (build with -ffp-contract=off). The CUDA product kernels are compiled with -fmad=false and
 * write the hit test / compositing expressions the same way, so hit ordering is reproducible
 * bit for bit; only expf/logf (library implementations differ by <= 2 ulp) are not.
 *
 * Deliberate deviations from the reference, all documented in DESIGN.md:
 *   - rsqrtf (approximate intrinsic, auxiliary.h:306) is evaluated as 1/sqrt.
 *   - ORC_ANALYTIC hit test (default): the two proxy triangles of a Gaussian are one analytic
 *     quad |u|,|v| <= f in the surfel frame; ORC_TRIANGLES keeps the literal two-triangle
 *     geometry (fp64 Moeller-Trumbore standing in for OptiX's closed-source intersector).
 *   - The "t < 0.2 stale slot" behaviour (forward.cu:214 precedes the slot reset at :218) is
 *     order-dependent in the reference whenever a ray has >= 16 hits in a round AND a hit
 *     closer than 0.2; here such hits are simply skipped (they still occupy k-buffer slots).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORC_DOUBLE
typedef double real;
#define R_SQRT sqrt
#define R_EXP exp
#define R_LOG log
#define R_FABS fabs
#define R_FMIN fmin
#define R_FMAX fmax
#else
typedef float real;
#define R_SQRT sqrtf
#define R_EXP expf
#define R_LOG logf
#define R_FABS fabsf
#define R_FMIN fminf
#define R_FMAX fmaxf
#endif

#define ORC_CHUNK 16             /* config.h:16  CHUNK_SIZE */
#define ORC_STEP_EPS 0.00001     /* config.h:17  STEP_EPSILON (a double literal there too) */
#define ORC_NCH 9                /* config.h:24  NUM_CHANNELS_F */

/* flags */
#define ORC_TRIANGLES 1          /* literal two-triangle proxy instead of the analytic quad */
#define ORC_BVH 2                /* candidate filter (median BVH) instead of testing every Gaussian */
#define ORC_FIX_BG 4             /* drop the duplicated background term of backward.cu:595-598 */
#define ORC_FLAT 8               /* ANALYSIS MODE, not the reference's arithmetic: the ray's hits are collected ONCE from the original
                                  * origin and walked in (t, id) order; a round is the next 16 of them with t > base, and the depth of a
                                  * hit is its t (the reference re-traces from o + base d each round and uses t' + base). Same rules
                                  * otherwise. Quantifies what a candidate-parallel evaluation of the hits would change. */

static const real SH_C0 = (real)0.28209479177387814;          /* auxiliary.h:23-40 */
static const real SH_C1 = (real)0.4886025119029199;
static const real SH_C2[5] = {(real)1.0925484305920792, (real)-1.0925484305920792, (real)0.31539156525252005,
                              (real)-1.0925484305920792, (real)0.5462742152960396};
static const real SH_C3[7] = {(real)-0.5900435899266435, (real)2.890611442640554, (real)-0.4570457994644658,
                              (real)0.3731763325901154, (real)-0.4570457994644658, (real)1.445305721320277,
                              (real)-0.5900435899266435};

typedef struct {
    real mu[3];
    real tu[3], tv[3], n[3];     /* columns of R (build_rotation) */
    real Lu[3], Lv[3];           /* rows 0,1 of L = S^-1 R^T (forward.cu:130-132) */
    real sx, sy, op, f;          /* f = cutoff (primitive_utils.py:200-201, backward.cu:625) */
    real qn[4];                  /* normalised quaternion (w,x,y,z) */
    real v[4][3];                /* quad corners, build2DRectangle order */
    real lo[3], hi[3];           /* padded AABB (candidate filter only) */
    int valid;
} gauss_t;

typedef struct {
    int P;
    gauss_t* g;
    /* Morton-ordered BVH over Gaussians (candidate filter; validated against brute force in tests) */
    int n_nodes;
    real* nlo; real* nhi;        /* 3 per node */
    int* nleft; int* nright; int* nfirst; int* ncount;
    int* order;
} scene_t;

typedef struct { real t; int id; } hit_t;   /* id = primitive (triangles) or Gaussian (analytic) */

/* ---------------------------------------------------------------------------------------- */

static void derive(const real* mu, const real* sc, const real* q, real op, real mod, gauss_t* o)
{
    /* general_utils.py:176-197 / auxiliary.h:306-328: normalise, then the standard matrix */
    const real nrm = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const real inv = (real)1 / R_SQRT(nrm);
    const real w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    o->qn[0] = w; o->qn[1] = x; o->qn[2] = y; o->qn[3] = z;
    const real R00 = (real)1 - (real)2 * (y * y + z * z), R01 = (real)2 * (x * y - w * z), R02 = (real)2 * (x * z + w * y);
    const real R10 = (real)2 * (x * y + w * z), R11 = (real)1 - (real)2 * (x * x + z * z), R12 = (real)2 * (y * z - w * x);
    const real R20 = (real)2 * (x * z - w * y), R21 = (real)2 * (y * z + w * x), R22 = (real)1 - (real)2 * (x * x + y * y);
    o->tu[0] = R00; o->tu[1] = R10; o->tu[2] = R20;
    o->tv[0] = R01; o->tv[1] = R11; o->tv[2] = R21;
    o->n[0] = R02; o->n[1] = R12; o->n[2] = R22;
    o->sx = sc[0]; o->sy = sc[1]; o->op = op;
    const real isx = (real)1 / (mod * sc[0]), isy = (real)1 / (mod * sc[1]);   /* auxiliary.h:445-452 */
    for (int k = 0; k < 3; k++) { o->mu[k] = mu[k]; o->Lu[k] = o->tu[k] * isx; o->Lv[k] = o->tv[k] * isy; }
    /* cutoff: sqrt(2 ln(255 o)) + 0.01 */
    o->f = R_SQRT((real)2 * R_LOG(op * (real)255)) + (real)0.01;
    o->valid = (o->f == o->f) && (o->f >= 0);
    /* corners: local (-1,1) (-1,-1) (1,1) (1,-1), world = R diag(sx f, sy f, 1) l + mu */
    static const real lx[4] = {-1, -1, 1, 1}, ly[4] = {1, -1, 1, -1};
    const real ax = sc[0] * o->f, ay = sc[1] * o->f;
    for (int c = 0; c < 4; c++)
        for (int k = 0; k < 3; k++)
            o->v[c][k] = (lx[c] * (o->tu[k] * ax) + ly[c] * (o->tv[k] * ay)) + mu[k];
    for (int k = 0; k < 3; k++) {
        /* candidate-filter box: must contain the analytic quad (half-extent mod*s*f) AND the reference triangles (s*f) */
        const real e = (R_FABS(o->tu[k]) * ax + R_FABS(o->tv[k]) * ay) * (mod > (real)1 ? mod : (real)1);
        const real pad = (real)1e-4 + (real)1e-5 * (R_FABS(mu[k]) + e);
        o->lo[k] = mu[k] - e - pad; o->hi[k] = mu[k] + e + pad;
    }
}

/* ---------------- candidate filter: Morton-ordered binary BVH over Gaussian AABBs ---------------- */

/* Morton order once (qsort), then split every range at its middle index; boxes bottom-up. */
typedef struct { uint64_t key; int id; } mkey_t;
static int cmp_mkey(const void* a, const void* b)
{
    const mkey_t* x = (const mkey_t*)a; const mkey_t* y = (const mkey_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return (x->id > y->id) - (x->id < y->id);
}
static uint64_t spread21(uint64_t v)
{
    v &= 0x1fffffULL;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}

static int bvh_rec(scene_t* s, int first, int count)
{
    const int id = s->n_nodes++;
    s->nfirst[id] = first; s->ncount[id] = 0; s->nleft[id] = s->nright[id] = -1;
    real* lo = s->nlo + 3 * id; real* hi = s->nhi + 3 * id;
    if (count <= 8) {
        for (int k = 0; k < 3; k++) { lo[k] = (real)1e30; hi[k] = (real)-1e30; }
        for (int i = first; i < first + count; i++) {
            const gauss_t* g = &s->g[s->order[i]];
            for (int k = 0; k < 3; k++) { if (g->lo[k] < lo[k]) lo[k] = g->lo[k]; if (g->hi[k] > hi[k]) hi[k] = g->hi[k]; }
        }
        s->ncount[id] = count;
        return id;
    }
    const int half = count / 2;
    const int l = bvh_rec(s, first, half);
    const int r = bvh_rec(s, first + half, count - half);
    s->nleft[id] = l; s->nright[id] = r;
    lo = s->nlo + 3 * id; hi = s->nhi + 3 * id;
    for (int k = 0; k < 3; k++) {
        lo[k] = s->nlo[3 * l + k] < s->nlo[3 * r + k] ? s->nlo[3 * l + k] : s->nlo[3 * r + k];
        hi[k] = s->nhi[3 * l + k] > s->nhi[3 * r + k] ? s->nhi[3 * l + k] : s->nhi[3 * r + k];
    }
    return id;
}

static void bvh_build(scene_t* s)
{
    int nv = 0;
    real clo[3] = {(real)1e30, (real)1e30, (real)1e30}, chi[3] = {(real)-1e30, (real)-1e30, (real)-1e30};
    mkey_t* mk = (mkey_t*)malloc(sizeof(mkey_t) * (size_t)(s->P > 0 ? s->P : 1));
    for (int i = 0; i < s->P; i++) if (s->g[i].valid)
        for (int k = 0; k < 3; k++) { if (s->g[i].mu[k] < clo[k]) clo[k] = s->g[i].mu[k]; if (s->g[i].mu[k] > chi[k]) chi[k] = s->g[i].mu[k]; }
    for (int i = 0; i < s->P; i++) if (s->g[i].valid) {
        uint64_t key = 0;
        for (int k = 0; k < 3; k++) {
            const double ext = (double)chi[k] - (double)clo[k];
            double t = ext > 0 ? ((double)s->g[i].mu[k] - (double)clo[k]) / ext * 2097151.0 : 0.0;
            if (t < 0) t = 0; if (t > 2097151.0) t = 2097151.0;
            key |= spread21((uint64_t)t) << (2 - k);
        }
        mk[nv].key = key; mk[nv].id = i; nv++;
    }
    qsort(mk, (size_t)nv, sizeof(mkey_t), cmp_mkey);
    s->order = (int*)malloc(sizeof(int) * (size_t)(nv > 0 ? nv : 1));
    for (int i = 0; i < nv; i++) s->order[i] = mk[i].id;
    free(mk);
    const int cap = 2 * (nv > 0 ? nv : 1);
    s->nlo = (real*)malloc(sizeof(real) * 3 * (size_t)cap); s->nhi = (real*)malloc(sizeof(real) * 3 * (size_t)cap);
    s->nleft = (int*)malloc(sizeof(int) * (size_t)cap); s->nright = (int*)malloc(sizeof(int) * (size_t)cap);
    s->nfirst = (int*)malloc(sizeof(int) * (size_t)cap); s->ncount = (int*)malloc(sizeof(int) * (size_t)cap);
    s->n_nodes = 0;
    if (nv > 0) bvh_rec(s, 0, nv);
}

static scene_t* scene_make(int P, const real* means, const real* scales, const real* rots, const real* opac,
                           real mod, int flags)
{
    scene_t* s = (scene_t*)calloc(1, sizeof(scene_t));
    s->P = P;
    s->g = (gauss_t*)malloc(sizeof(gauss_t) * (size_t)(P > 0 ? P : 1));
    for (int i = 0; i < P; i++) derive(means + 3 * i, scales + 2 * i, rots + 4 * i, opac[i], mod, &s->g[i]);
    if (flags & ORC_BVH) bvh_build(s);
    return s;
}

static void scene_free(scene_t* s)
{
    free(s->g); free(s->nlo); free(s->nhi); free(s->nleft); free(s->nright); free(s->nfirst); free(s->ncount);
    free(s->order); free(s);
}

/* --------------------------------- hit tests ------------------------------------------- */

/* analytic quad: t' on the surfel plane from the (re-based) origin, then |u|,|v| <= f */
static inline int quad_hit(const gauss_t* g, const real* o, const real* d, real* t_out)
{
    const real c0 = g->mu[0] - o[0], c1 = g->mu[1] - o[1], c2 = g->mu[2] - o[2];
    const real den = g->n[0] * d[0] + g->n[1] * d[1] + g->n[2] * d[2];
    const real num = g->n[0] * c0 + g->n[1] * c1 + g->n[2] * c2;
    const real t = num / den;
    if (!(t > 0)) return 0;
    const real r0 = (o[0] + t * d[0]) - g->mu[0], r1 = (o[1] + t * d[1]) - g->mu[1], r2 = (o[2] + t * d[2]) - g->mu[2];
    const real u = g->Lu[0] * r0 + g->Lu[1] * r1 + g->Lu[2] * r2;
    const real v = g->Lv[0] * r0 + g->Lv[1] * r1 + g->Lv[2] * r2;
    if (!(R_FABS(u) <= g->f && R_FABS(v) <= g->f)) return 0;
    *t_out = t;
    return 1;
}

/* literal triangle (fp64 Moeller-Trumbore standing in for OptiX's intersector) */
static inline double tri_hit(const real* o, const real* d, const real* A, const real* B, const real* C)
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

typedef struct { hit_t* h; int n, cap; } hitbuf_t;

static inline void hb_push(hitbuf_t* b, real t, int id)
{
    if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 256; b->h = (hit_t*)realloc(b->h, sizeof(hit_t) * (size_t)b->cap); }
    b->h[b->n].t = t; b->h[b->n].id = id; b->n++;
}

static inline void test_gauss(const scene_t* s, int gi, const real* o, const real* d, int flags, hitbuf_t* hb)
{
    const gauss_t* g = &s->g[gi];
    if (!g->valid) return;
    if (flags & ORC_TRIANGLES) {
        /* prim 2g = (v0,v1,v2), prim 2g+1 = (v2,v3,v1): primitive_utils.py:212-221; tmax 1e16: forward.cu:55 */
        const double ta = tri_hit(o, d, g->v[0], g->v[1], g->v[2]);
        if (ta > 0.0 && (float)ta > 0.0f && (float)ta < 1e16f) hb_push(hb, (real)(float)ta, 2 * gi);
        const double tb = tri_hit(o, d, g->v[2], g->v[3], g->v[1]);
        if (tb > 0.0 && (float)tb > 0.0f && (float)tb < 1e16f) hb_push(hb, (real)(float)tb, 2 * gi + 1);
    } else {
        real t;
        if (quad_hit(g, o, d, &t) && t < (real)1e16) hb_push(hb, t, gi);
    }
}

static inline int ray_box(const real* lo, const real* hi, const real* o, const real* inv)
{
    real t0 = 0, t1 = (real)1e30;
    for (int k = 0; k < 3; k++) {
        real a = (lo[k] - o[k]) * inv[k], c = (hi[k] - o[k]) * inv[k];
        if (a != a || c != c) continue;
        if (a > c) { const real tmp = a; a = c; c = tmp; }
        if (a > t0) t0 = a;
        if (c < t1) t1 = c;
    }
    return t0 <= t1 * (real)1.00001 + (real)1e-6;
}

static int cmp_hit(const void* a, const void* b)
{
    const hit_t* x = (const hit_t*)a; const hit_t* y = (const hit_t*)b;
    if (x->t < y->t) return -1;
    if (x->t > y->t) return 1;
    return (x->id > y->id) - (x->id < y->id);
}

/* every proxy hit of ray (o, d) with t' > 0, ascending (t', id) — what the any-hit k-buffer of
 * forward.cu:312-356 converges to for its first 16 entries, regardless of traversal order */
static void collect_hits(const scene_t* s, const real* o, const real* d, int flags, hitbuf_t* hb)
{
    hb->n = 0;
    if ((flags & ORC_BVH) && s->n_nodes > 0) {
        const real inv[3] = {(real)1 / d[0], (real)1 / d[1], (real)1 / d[2]};
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const int id = stack[--sp];
            if (!ray_box(s->nlo + 3 * id, s->nhi + 3 * id, o, inv)) continue;
            if (s->ncount[id] > 0) {
                for (int i = s->nfirst[id]; i < s->nfirst[id] + s->ncount[id]; i++) test_gauss(s, s->order[i], o, d, flags, hb);
            } else { stack[sp++] = s->nleft[id]; stack[sp++] = s->nright[id]; }
        }
    } else if (!(flags & ORC_BVH)) {
        for (int i = 0; i < s->P; i++) test_gauss(s, i, o, d, flags, hb);
    }
    if (hb->n > 1) qsort(hb->h, (size_t)hb->n, sizeof(hit_t), cmp_hit);
}

/* ------------------------------------ SH ------------------------------------------------ */

/* forward.cu:67-111 — basis(dir) . sh + 0.5, channel 0 clamped at 0. basis[] returned for the VJP. */
static void sh_eval(int deg, const real* dirn, const real* sh /* [M][3] */, real* c, int* clamped0, real* basis)
{
    const real x = dirn[0], y = dirn[1], z = dirn[2];
    int nb = 1;
    basis[0] = SH_C0;
    if (deg > 0) {
        basis[1] = -SH_C1 * y; basis[2] = SH_C1 * z; basis[3] = -SH_C1 * x; nb = 4;
        if (deg > 1) {
            const real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            basis[4] = SH_C2[0] * xy; basis[5] = SH_C2[1] * yz; basis[6] = SH_C2[2] * ((real)2 * zz - xx - yy);
            basis[7] = SH_C2[3] * xz; basis[8] = SH_C2[4] * (xx - yy); nb = 9;
            if (deg > 2) {
                basis[9] = SH_C3[0] * y * ((real)3 * xx - yy);
                basis[10] = SH_C3[1] * xy * z;
                basis[11] = SH_C3[2] * y * ((real)4 * zz - xx - yy);
                basis[12] = SH_C3[3] * z * ((real)2 * zz - (real)3 * xx - (real)3 * yy);
                basis[13] = SH_C3[4] * x * ((real)4 * zz - xx - yy);
                basis[14] = SH_C3[5] * z * (xx - yy);
                basis[15] = SH_C3[6] * x * (xx - (real)3 * yy);
                nb = 16;
            }
        }
    }
    for (int ch = 0; ch < 3; ch++) {
        real r = basis[0] * sh[ch];
        for (int j = 1; j < nb; j++) r = r + basis[j] * sh[3 * j + ch];
        c[ch] = r + (real)0.5;
    }
    *clamped0 = c[0] < 0;
    if (c[0] < 0) c[0] = 0;
    for (int j = nb; j < 16; j++) basis[j] = 0;
}

/* ------------------------------- per-ray program ---------------------------------------- */

typedef struct {
    /* inputs */
    const scene_t* s; int flags; int D, M; const real* shs; const real* bg;
    /* forward outputs */
    real* out; real* accum_w; int* hit_list; int* hit_cnt; int cap; int* slot_cnt;
    /* backward (NULL in forward) */
    const real* fwd_out; const real* dL_dout;
    real* d_means; real* d_shs; real* d_opac; real* d_scales; real* d_rots;
} job_t;

static inline void atomic_add(real* p, real v)
{
#pragma omp atomic
    *p += v;
}

/* auxiliary.h:389-433. G[r][c] = dL/dR(r,c) for the standard matrix R of the normalised q. */
static void quat_vjp(const real* qn, real G[3][3], real* dq)
{
    const real w = qn[0], x = qn[1], y = qn[2], z = qn[3];
    dq[0] = (real)2 * (x * (G[2][1] - G[1][2]) + y * (G[0][2] - G[2][0]) + z * (G[1][0] - G[0][1]));
    dq[1] = (real)2 * ((real)-2 * x * (G[1][1] + G[2][2]) + y * (G[1][0] + G[0][1]) + z * (G[2][0] + G[0][2]) + w * (G[2][1] - G[1][2]));
    dq[2] = (real)2 * (x * (G[1][0] + G[0][1]) - (real)2 * y * (G[0][0] + G[2][2]) + z * (G[2][1] + G[1][2]) + w * (G[0][2] - G[2][0]));
    dq[3] = (real)2 * (x * (G[2][0] + G[0][2]) + y * (G[2][1] + G[1][2]) - (real)2 * z * (G[0][0] + G[1][1]) + w * (G[1][0] - G[0][1]));
}

static inline void cross3(const real* a, const real* b, real* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

static void ray_program(const job_t* J, int r, const real* ro, const real* rd, hitbuf_t* hb)
{
    const scene_t* s = J->s;
    const int bw = J->dL_dout != NULL;
    real C[3] = {0, 0, 0}, N[3] = {0, 0, 0}, Dp = 0, W = 0, T = 1, testT = 1;
    real base = 0, dpt = 0;
    int last = -1, ncontrib = 0, nslots = 0;
    const real dl = R_SQRT(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
    const real dirn[3] = {rd[0] / dl, rd[1] / dl, rd[2] / dl};
    real g_rgb[3] = {0, 0, 0}, g_d = 0, g_n[3] = {0, 0, 0}, F_c[3] = {0, 0, 0}, F_d = 0, F_n[3] = {0, 0, 0}, F_T = 0;
    if (bw) {
        /* backward.cu:472-493 (dL_daccum and dL_dT are read/ignored there) */
        for (int c = 0; c < 3; c++) { g_rgb[c] = J->dL_dout[ORC_NCH * r + c]; g_n[c] = J->dL_dout[ORC_NCH * r + 5 + c]; }
        g_d = J->dL_dout[ORC_NCH * r + 3];
        for (int c = 0; c < 3; c++) { F_c[c] = J->fwd_out[ORC_NCH * r + c]; F_n[c] = J->fwd_out[ORC_NCH * r + 5 + c]; }
        F_d = J->fwd_out[ORC_NCH * r + 3]; F_T = J->fwd_out[ORC_NCH * r + 8];
    }
    const int flat = (J->flags & ORC_FLAT) != 0;
    int fpos = 0;                                    /* ORC_FLAT: next unread hit of the one list */
    if (flat) collect_hits(s, ro, rd, J->flags, hb);
    for (;;) {
        const real om[3] = {ro[0] + base * rd[0], ro[1] + base * rd[1], ro[2] + base * rd[2]};   /* forward.cu:291 */
        int cnt, first = 0;
        if (!flat) {
            collect_hits(s, om, rd, J->flags, hb);
            cnt = hb->n;                             /* >= 16 iff the k-buffer filled */
        } else {
            while (fpos < hb->n && !(hb->h[fpos].t > base)) fpos++;      /* t' = t - base > 0 */
            first = fpos; cnt = hb->n - fpos;
        }
        const int lim = cnt < ORC_CHUNK ? cnt : ORC_CHUNK;
        if (flat) fpos += lim;
        int terminated = 0;
        for (int i = first; i < first + lim; i++) {
            const int pid = hb->h[i].id;
            const int gi = (J->flags & ORC_TRIANGLES) ? pid / 2 : pid;
            const gauss_t* g = &s->g[gi];
            nslots++;
            dpt = flat ? hb->h[i].t : hb->h[i].t + base;                 /* forward.cu:212 */
            const real xyz[3] = {ro[0] + dpt * rd[0], ro[1] + dpt * rd[1], ro[2] + dpt * rd[2]};
            if (dpt < (real)0.2) continue;                               /* :214 */
            if (gi == last) continue;                                    /* :220-224 */
            last = gi;
            const real r0 = xyz[0] - g->mu[0], r1 = xyz[1] - g->mu[1], r2 = xyz[2] - g->mu[2];
            const real u = g->Lu[0] * r0 + g->Lu[1] * r1 + g->Lu[2] * r2;   /* :139 */
            const real v = g->Lv[0] * r0 + g->Lv[1] * r1 + g->Lv[2] * r2;
            const real cosv = -((g->mu[0] - ro[0]) * g->n[0] + (g->mu[1] - ro[1]) * g->n[1] + (g->mu[2] - ro[2]) * g->n[2]);
            if (!bw && cosv == 0) continue;                              /* :233-237 (forward only) */
            const real rho = u * u + v * v;
            const real power = (real)-0.5 * rho;
            if (power > 0) continue;
            const real G = R_EXP(power);
            const real alpha = R_FMIN((real)0.99, g->op * G);            /* :249 */
            if (alpha < (real)1 / (real)255) continue;
            testT = T * ((real)1 - alpha);
            if (testT < (real)0.0001) { terminated = 1; break; }         /* :253-257 */
            const real w = alpha * T;
            real c[3], basis[16]; int clamped0;
            sh_eval(J->D, dirn, J->shs + (size_t)gi * J->M * 3, c, &clamped0, basis);
            for (int ch = 0; ch < 3; ch++) C[ch] += w * c[ch];
            Dp += w * dpt; W += w;
            if (!bw) {
                atomic_add(&J->accum_w[gi], w);                          /* :272 */
                if (J->hit_list && ncontrib < J->cap) J->hit_list[(size_t)r * J->cap + ncontrib] = gi;
            } else {
                /* ---------------- backward.cu:577-675 ---------------- */
                for (int k = 0; k < 3; k++) N[k] += w * g->n[k];         /* :578 (not sign-adjusted there) */
                const real inv1a = (real)1 / ((real)1 - alpha);
                real dalpha = 0, dcol[3];
                for (int ch = 0; ch < 3; ch++) {
                    dcol[ch] = g_rgb[ch] * w;
                    dalpha += g_rgb[ch] * (T * c[ch] - (F_c[ch] - C[ch]) * inv1a);
                }
                if (!(J->flags & ORC_FIX_BG)) {
                    real dbg = 0;
                    for (int ch = 0; ch < 3; ch++) dbg += g_rgb[ch] * J->bg[ch];
                    dalpha += dbg * (-F_T * inv1a);                      /* :595-598 */
                }
                const real dD_gs = g_d * w;
                dalpha += g_d * (T * dpt - (F_d - Dp) * inv1a);          /* :601 */
                real dN_gs[3];
                {
                    real acc = 0;
                    for (int k = 0; k < 3; k++) { dN_gs[k] = g_n[k] * w; acc += g_n[k] * (T * g->n[k] - (F_n[k] - N[k]) * inv1a); }
                    dalpha += acc;                                       /* :604 */
                }
                if (g->op * G > (real)0.99) dalpha = 0;                  /* :607-608 */
                const real dG = g->op * dalpha;
                atomic_add(&J->d_opac[gi], G * dalpha);                  /* :615 */
                const real nsign = cosv > 0 ? (real)1 : (real)-1;        /* :649-650 */
                /* compute_transmat_uv_backward, backward.cu:339-431 */
                const real du = dG * -G * u, dv = dG * -G * v;
                real dtu[3], dtv[3], dn[3], dmu[3], dsc[2];
                const real rr[3] = {r0, r1, r2};
                for (int k = 0; k < 3; k++) { dtu[k] = du * rr[k] / g->sx; dtv[k] = dv * rr[k] / g->sy; dn[k] = dN_gs[k] * nsign; }
                dsc[0] = dG * (G * u * u / g->sx); dsc[1] = dG * (G * v * v / g->sy);
                for (int k = 0; k < 3; k++) dmu[k] = dG * (G * (g->Lu[k] * u + g->Lv[k] * v));
                real dxyz[3];
                for (int k = 0; k < 3; k++) dxyz[k] = du * g->Lu[k] + dv * g->Lv[k];
                const real dd = (dxyz[0] * rd[0] + dxyz[1] * rd[1] + dxyz[2] * rd[2]) + dD_gs;   /* :390 */
                /* hit triangle (backward.cu:628-647): even -> verts 0,1,2 ; odd -> verts 1,2,3 */
                int odd;
                if (J->flags & ORC_TRIANGLES) odd = pid & 1;
                else odd = !(v >= u);      /* (-1,1) corner side of the (-1,-1)-(1,1) diagonal */
                const real* v1 = g->v[odd ? 1 : 0]; const real* v2 = g->v[odd ? 2 : 1]; const real* v3 = g->v[odd ? 3 : 2];
                /* cutoff recomputed in double there: sqrt(2.0f * log(opacity * 255.)) + 0.01 */
                const real cutoff = (real)(sqrt(2.0 * log((double)g->op * 255.)) + 0.01);
                static const real hx[4] = {-1, -1, 1, 1}, hy[4] = {1, -1, 1, -1};
                const int i1 = odd ? 1 : 0, i2 = odd ? 2 : 1, i3 = odd ? 3 : 2;
                const real h1x = hx[i1] * cutoff, h1y = hy[i1] * cutoff, h2x = hx[i2] * cutoff, h2y = hy[i2] * cutoff,
                           h3x = hx[i3] * cutoff, h3y = hy[i3] * cutoff;
                real e21[3], e31[3], nT[3], cT[3];
                for (int k = 0; k < 3; k++) { e21[k] = v2[k] - v1[k]; e31[k] = v3[k] - v1[k]; cT[k] = v1[k] - ro[k]; }
                cross3(e21, e31, nT);
                const real pp = nT[0] * cT[0] + nT[1] * cT[1] + nT[2] * cT[2];
                const real qq = nT[0] * rd[0] + nT[1] * rd[1] + nT[2] * rd[2];
                real a[3];
                for (int k = 0; k < 3; k++) a[k] = (cT[k] - pp / qq * rd[k]) / qq;
                real e23[3], e12[3], x1[3], x2[3], x3[3], dv1[3], dv2[3], dv3[3];
                for (int k = 0; k < 3; k++) { e23[k] = v2[k] - v3[k]; e12[k] = v1[k] - v2[k]; }
                cross3(e23, a, x1); cross3(e31, a, x2); cross3(e12, a, x3);
                for (int k = 0; k < 3; k++) { dv1[k] = x1[k] * dd + nT[k] / qq * dd; dv2[k] = x2[k] * dd; dv3[k] = x3[k] * dd; }
                real Sx[3], Sy[3];
                for (int k = 0; k < 3; k++) {
                    Sx[k] = h1x * dv1[k] + h2x * dv2[k] + h3x * dv3[k];
                    Sy[k] = h1y * dv1[k] + h2y * dv2[k] + h3y * dv3[k];
                    dtu[k] += g->sx * Sx[k]; dtv[k] += g->sy * Sy[k];
                }
                dsc[0] += (g->sx * g->Lu[0]) * Sx[0] + (g->sx * g->Lu[1]) * Sx[1] + (g->sx * g->Lu[2]) * Sx[2];   /* :426 */
                dsc[1] += (g->sy * g->Lv[0]) * Sy[0] + (g->sy * g->Lv[1]) * Sy[1] + (g->sy * g->Lv[2]) * Sy[2];
                for (int k = 0; k < 3; k++) dmu[k] += dv1[k] + dv2[k] + dv3[k];
                real Gm[3][3], dq[4];
                for (int k = 0; k < 3; k++) { Gm[k][0] = dtu[k]; Gm[k][1] = dtv[k]; Gm[k][2] = dn[k]; }
                quat_vjp(g->qn, Gm, dq);
                atomic_add(&J->d_scales[2 * gi], dsc[0]); atomic_add(&J->d_scales[2 * gi + 1], dsc[1]);
                for (int k = 0; k < 4; k++) atomic_add(&J->d_rots[4 * gi + k], dq[k]);
                for (int k = 0; k < 3; k++) atomic_add(&J->d_means[3 * gi + k], dmu[k]);
                /* computeColorFromSHBackward, backward.cu:123-247 */
                if (clamped0) dcol[0] = 0;
                const int nb = (J->D + 1) * (J->D + 1);
                for (int j = 0; j < nb; j++)
                    for (int ch = 0; ch < 3; ch++) atomic_add(&J->d_shs[((size_t)gi * J->M + j) * 3 + ch], basis[j] * dcol[ch]);
            }
            ncontrib++;
            T = testT;
        }
        if (terminated || testT < (real)0.0001 || cnt < ORC_CHUNK) break;                  /* forward.cu:282-285 */
        base = (real)((double)dpt + ORC_STEP_EPS);                                          /* :288 */
    }
    if (!bw) {
        real* o = J->out + (size_t)ORC_NCH * r;
        for (int ch = 0; ch < 3; ch++) o[ch] = C[ch] + T * J->bg[ch];    /* :296-305 */
        o[3] = Dp; o[4] = W; o[5] = 0; o[6] = 0; o[7] = 0; o[8] = T;
        if (J->hit_cnt) J->hit_cnt[r] = ncontrib;
        if (J->slot_cnt) J->slot_cnt[r] = nslots;
    }
}

static void run(job_t* J, int R, const real* ray_o, int ray_o_stride, const real* ray_d)
{
#pragma omp parallel
    {
        hitbuf_t hb = {0, 0, 0};
#pragma omp for schedule(dynamic, 64)
        for (int r = 0; r < R; r++) ray_program(J, r, ray_o + (size_t)r * ray_o_stride, ray_d + (size_t)3 * r, &hb);
        free(hb.h);
    }
}

/* ------------------------------------ C API --------------------------------------------- */

/* forward: out (R,9), accum_w (P); optional hit_list (R,cap) int32 = contributing Gaussian ids in
 * compositing order, hit_cnt (R) = number of contributing hits (may exceed cap), slot_cnt (R) =
 * k-buffer slots consumed. ray_o_stride: 3 (per-ray origins) or 0 (one shared origin). */
int orc_forward(int R, const real* ray_o, int ray_o_stride, const real* ray_d, const real* bg,
                int P, const real* means, const real* scales, const real* rots, const real* opac,
                const real* shs, int D, int M, real scale_modifier, int flags,
                real* out, real* accum_w, int* hit_list, int* hit_cnt, int cap, int* slot_cnt)
{
    scene_t* s = scene_make(P, means, scales, rots, opac, scale_modifier, flags);
    memset(accum_w, 0, sizeof(real) * (size_t)P);
    job_t J; memset(&J, 0, sizeof(J));
    J.s = s; J.flags = flags; J.D = D; J.M = M; J.shs = shs; J.bg = bg;
    J.out = out; J.accum_w = accum_w; J.hit_list = hit_list; J.hit_cnt = hit_cnt; J.cap = cap; J.slot_cnt = slot_cnt;
    run(&J, R, ray_o, ray_o_stride, ray_d);
    scene_free(s);
    return 0;
}

int orc_backward(int R, const real* ray_o, int ray_o_stride, const real* ray_d, const real* bg,
                 int P, const real* means, const real* scales, const real* rots, const real* opac,
                 const real* shs, int D, int M, real scale_modifier, int flags,
                 const real* fwd_out, const real* dL_dout,
                 real* d_means, real* d_shs, real* d_opac, real* d_scales, real* d_rots)
{
    scene_t* s = scene_make(P, means, scales, rots, opac, scale_modifier, flags);
    memset(d_means, 0, sizeof(real) * (size_t)P * 3); memset(d_shs, 0, sizeof(real) * (size_t)P * M * 3);
    memset(d_opac, 0, sizeof(real) * (size_t)P); memset(d_scales, 0, sizeof(real) * (size_t)P * 2);
    memset(d_rots, 0, sizeof(real) * (size_t)P * 4);
    job_t J; memset(&J, 0, sizeof(J));
    J.s = s; J.flags = flags; J.D = D; J.M = M; J.shs = shs; J.bg = bg;
    J.fwd_out = fwd_out; J.dL_dout = dL_dout;
    J.d_means = d_means; J.d_shs = d_shs; J.d_opac = d_opac; J.d_scales = d_scales; J.d_rots = d_rots;
    run(&J, R, ray_o, ray_o_stride, ray_d);
    scene_free(s);
    return 0;
}

/* build2DRectangle (primitive_utils.py:182-224): vertices (4P,3). faces are implicit:
 * (4g,4g+1,4g+2) and (4g+2,4g+3,4g+1). */
int orc_build_rectangles(int P, const real* means, const real* scales, const real* rots, const real* opac, real* verts)
{
    for (int i = 0; i < P; i++) {
        gauss_t g; derive(means + 3 * i, scales + 2 * i, rots + 4 * i, opac[i], (real)1, &g);
        for (int c = 0; c < 4; c++) for (int k = 0; k < 3; k++) verts[((size_t)4 * i + c) * 3 + k] = g.v[c][k];
    }
    return 0;
}

/* SH colour of one direction (for the eval_sh golden vectors): c[3], basis[16] */
int orc_sh_eval(int deg, const real* dir, const real* sh, real* c, real* basis)
{
    const real dl = R_SQRT(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    const real dn[3] = {dir[0] / dl, dir[1] / dl, dir[2] / dl};
    int cl;
    sh_eval(deg, dn, sh, c, &cl, basis);
    return cl;
}

int orc_real_size(void) { return (int)sizeof(real); }
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
