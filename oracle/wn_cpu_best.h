/* TEST / BENCH INFRASTRUCTURE -- not part of the product.
 *
 * wn_cpu_best.h: the "best-effort CPU" baseline of BASELINE.md section 3 (B-cpu) for the WaveNet sample loop
 * (generate.py:202-233 -> wavenet/model.py:112-167, 215-245), included at the end of wn_oracle.c so that it shares the
 * oracle's model structures, pinned math (wn_math_ref.h) and draws.  It is the CPU analogue of the GPU design:
 *
 *   - weight-stationary threads: thread i owns a fixed column slice of every matrix (filter/gate, dense, skip, conv1), so
 *     its 1/nt of the weights stays in its private L2 from step to step instead of being re-streamed per row;
 *   - batch-major: all N rows advance together, every weight vector that is loaded is used for N rows;
 *   - the dependent chain is cut at the same places as on the GPU: 2 spin barriers per layer (z complete, x complete) and
 *     3 for the post-processing stack per step; the draw runs one row per thread.
 *
 * Arithmetic: the per-column operation sequence is exactly mv_plan's (t contiguous fma chains per column, ascending
 * xor-butterfly, then the oracle's additions in the oracle's order), so the output is BIT-IDENTICAL to orc_generate for the same
 * plan -- tests/test_oracle.py checks that; nothing is traded for speed except the order in which independent columns and
 * rows are visited.  bench.py reports it as cpu_baseline.best_effort next to the plain port. */
#include <stdatomic.h>
#include <immintrin.h>

typedef struct {
    atomic_int count;
    atomic_int sense;
    int n;
} orcb_barrier;

static inline void orcb_wait(orcb_barrier *b, int *local_sense)
{
    const int s = !*local_sense;
    *local_sense = s;
    if (atomic_fetch_add_explicit(&b->count, 1, memory_order_acq_rel) == b->n - 1) {
        atomic_store_explicit(&b->count, 0, memory_order_relaxed);
        atomic_store_explicit(&b->sense, s, memory_order_release);
    } else {
        while (atomic_load_explicit(&b->sense, memory_order_acquire) != s) _mm_pause();
    }
}

/* mv_plan for the column slice [c0, c1) of W and N input rows at once: out[b*ldo + o] for o in [c0, c1).
 * scratch: N * t * (c1 - c0) floats.  Per column the operations are those of mv_plan, in the same order; the loops run
 * chunk-outer / row-inner so that a chunk's weights (K/t x nc floats) are read from L1 for every row after the first, and the
 * accumulators of one (chunk, row) stay in registers when nc is one of the specialised widths. */
/* One chunk (ch consecutive k) of a 16- or 8-column strip for FOUR rows: 4 x 2 (or 4 x 1) independent fma chains held in
 * registers.  _mm256_fmadd_ps(w, x, a) is fmaf(w, x, a) per lane, so every column's chain is the oracle's, bit for bit.  (Written
 * with intrinsics because gcc turned the equivalent array loops into stack-resident accumulators once nothing at the call site
 * was a compile-time constant: 1.3 instead of 14 GFMA/s per core.) */
static inline void orcb_chain16x4(const float *wc, int stride, const float *x, int ldx, int ch, float *dst, size_t rs)
{
    __m256 a00 = _mm256_setzero_ps(), a01 = a00, a10 = a00, a11 = a00, a20 = a00, a21 = a00, a30 = a00, a31 = a00;
    const float *x0 = x, *x1 = x + ldx, *x2 = x1 + ldx, *x3 = x2 + ldx;
    for (int i = 0; i < ch; ++i) {
        const float *w = wc + (size_t)i * stride;
        const __m256 w0 = _mm256_loadu_ps(w), w1 = _mm256_loadu_ps(w + 8);
        const __m256 k0 = _mm256_broadcast_ss(x0 + i), k1 = _mm256_broadcast_ss(x1 + i), k2 = _mm256_broadcast_ss(x2 + i),
                     k3 = _mm256_broadcast_ss(x3 + i);
        a00 = _mm256_fmadd_ps(w0, k0, a00); a01 = _mm256_fmadd_ps(w1, k0, a01);
        a10 = _mm256_fmadd_ps(w0, k1, a10); a11 = _mm256_fmadd_ps(w1, k1, a11);
        a20 = _mm256_fmadd_ps(w0, k2, a20); a21 = _mm256_fmadd_ps(w1, k2, a21);
        a30 = _mm256_fmadd_ps(w0, k3, a30); a31 = _mm256_fmadd_ps(w1, k3, a31);
    }
    _mm256_storeu_ps(dst, a00); _mm256_storeu_ps(dst + 8, a01);
    _mm256_storeu_ps(dst + rs, a10); _mm256_storeu_ps(dst + rs + 8, a11);
    _mm256_storeu_ps(dst + 2 * rs, a20); _mm256_storeu_ps(dst + 2 * rs + 8, a21);
    _mm256_storeu_ps(dst + 3 * rs, a30); _mm256_storeu_ps(dst + 3 * rs + 8, a31);
}
static inline void orcb_chain8x4(const float *wc, int stride, const float *x, int ldx, int ch, float *dst, size_t rs)
{
    __m256 a0 = _mm256_setzero_ps(), a1 = a0, a2 = a0, a3 = a0;
    const float *x0 = x, *x1 = x + ldx, *x2 = x1 + ldx, *x3 = x2 + ldx;
    for (int i = 0; i < ch; ++i) {
        const __m256 w0 = _mm256_loadu_ps(wc + (size_t)i * stride);
        a0 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x0 + i), a0);
        a1 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x1 + i), a1);
        a2 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x2 + i), a2);
        a3 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x3 + i), a3);
    }
    _mm256_storeu_ps(dst, a0); _mm256_storeu_ps(dst + rs, a1); _mm256_storeu_ps(dst + 2 * rs, a2); _mm256_storeu_ps(dst + 3 * rs, a3);
}

/* eight rows of an 8-column strip: 8 independent chains (with four, the 4-cycle fma latency leaves half of the issue slots empty) */
static inline void orcb_chain8x8(const float *wc, int stride, const float *x, int ldx, int ch, float *dst, size_t rs)
{
    __m256 a0 = _mm256_setzero_ps(), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
    const float *x0 = x, *x1 = x + ldx, *x2 = x1 + ldx, *x3 = x2 + ldx, *x4 = x3 + ldx, *x5 = x4 + ldx, *x6 = x5 + ldx, *x7 = x6 + ldx;
    for (int i = 0; i < ch; ++i) {
        const __m256 w0 = _mm256_loadu_ps(wc + (size_t)i * stride);
        a0 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x0 + i), a0);
        a1 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x1 + i), a1);
        a2 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x2 + i), a2);
        a3 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x3 + i), a3);
        a4 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x4 + i), a4);
        a5 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x5 + i), a5);
        a6 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x6 + i), a6);
        a7 = _mm256_fmadd_ps(w0, _mm256_broadcast_ss(x7 + i), a7);
    }
    _mm256_storeu_ps(dst, a0); _mm256_storeu_ps(dst + rs, a1); _mm256_storeu_ps(dst + 2 * rs, a2); _mm256_storeu_ps(dst + 3 * rs, a3);
    _mm256_storeu_ps(dst + 4 * rs, a4); _mm256_storeu_ps(dst + 5 * rs, a5); _mm256_storeu_ps(dst + 6 * rs, a6); _mm256_storeu_ps(dst + 7 * rs, a7);
}

/* scratch block of a strip: [N][t][nc] */
static void mvb_chains(const float *W, int stride, const float *X, int ldx, int N, int K, int t, float *scratch, int nc)
{
    const int ch = K / t;
    const size_t rs = (size_t)t * nc;
    for (int c = 0; c < t; ++c) {
        const float *wc = W + (size_t)c * ch * stride;
        int b = 0;
        if (nc == 16)
            for (; b + 4 <= N; b += 4) orcb_chain16x4(wc, stride, X + (size_t)b * ldx + c * ch, ldx, ch, scratch + ((size_t)b * t + c) * nc, rs);
        else if (nc == 8) {
            for (; b + 8 <= N; b += 8) orcb_chain8x8(wc, stride, X + (size_t)b * ldx + c * ch, ldx, ch, scratch + ((size_t)b * t + c) * nc, rs);
            for (; b + 4 <= N; b += 4) orcb_chain8x4(wc, stride, X + (size_t)b * ldx + c * ch, ldx, ch, scratch + ((size_t)b * t + c) * nc, rs);
        }
        for (; b < N; ++b) {
            const float *x = X + (size_t)b * ldx + c * ch;
            float *restrict a = scratch + ((size_t)b * t + c) * nc;
            for (int o = 0; o < nc; ++o) a[o] = 0.0f;
            for (int i = 0; i < ch; ++i) {
                const float xk = x[i];
                const float *restrict w = wc + (size_t)i * stride;
                for (int o = 0; o < nc; ++o) a[o] = fmaf(w[o], xk, a[o]);
            }
        }
    }
}

static void mvb(const float *W, int stride, int c0, int c1, const float *X, int ldx, int N, int K, int t, float *out, int ldo,
                float *scratch)
{
    const int nc = c1 - c0;
    if (nc <= 0) return;
    /* column strips of 16 (two 256-bit or one 512-bit vector per row: RB x 2 accumulators), then 8, then the rest */
    int o0 = 0;
    for (; o0 + 16 <= nc; o0 += 16) mvb_chains(W + c0 + o0, stride, X, ldx, N, K, t, scratch + (size_t)o0 * N * t, 16);
    for (; o0 + 8 <= nc; o0 += 8) mvb_chains(W + c0 + o0, stride, X, ldx, N, K, t, scratch + (size_t)o0 * N * t, 8);
    if (o0 < nc) mvb_chains(W + c0 + o0, stride, X, ldx, N, K, t, scratch + (size_t)o0 * N * t, nc - o0);
    /* scratch: per strip (width sw) a block [N][t][sw] */
    for (o0 = 0; o0 < nc;) {
        const int sw = (nc - o0 >= 16) ? 16 : (nc - o0 >= 8) ? 8 : nc - o0;
        float *ss = scratch + (size_t)o0 * N * t;
        for (int b = 0; b < N; ++b) {
            float *sb = ss + (size_t)b * t * sw;
            for (int off = 1; off < t; off <<= 1)
                for (int c = 0; c < t; ++c)
                    if ((c & off) == 0) {
                        float *restrict a = sb + (size_t)c * sw;
                        float *restrict bb = sb + (size_t)(c ^ off) * sw;
                        for (int o = 0; o < sw; ++o) { float v = a[o] + bb[o]; a[o] = v; bb[o] = v; }
                    }
            float *restrict dst = out + (size_t)b * ldo + c0 + o0;
            for (int o = 0; o < sw; ++o) dst[o] = sb[o];
        }
        o0 += sw;
    }
}

/* The pinned activation (wn_math_ref.h: orc_exp32 / orc_tanh32 / orc_sigmoid32) for eight lanes at once: the same IEEE operations
 * in the same order per lane (fma, correctly rounded division, exponent add), so z = tanh32(f) * sigmoid32(g) is bit-identical to
 * the scalar functions -- tests/test_oracle.py compares them on edge values and random inputs. */
static inline __m256 orcb_exp32x8(__m256 x)
{
    const __m256 lt = _mm256_cmp_ps(x, _mm256_set1_ps(-87.0f), _CMP_LT_OQ);                   /* x < -87 -> +0 */
    x = _mm256_blendv_ps(x, _mm256_set1_ps(88.0f), _mm256_cmp_ps(x, _mm256_set1_ps(88.0f), _CMP_GT_OQ));
    const __m256 magic = _mm256_set1_ps(12582912.0f);
    const __m256 t = _mm256_fmadd_ps(x, _mm256_set1_ps(1.44269504088896341f), magic);
    const __m256 n = _mm256_sub_ps(t, magic);
    __m256 r = _mm256_fmadd_ps(n, _mm256_set1_ps(-0.693359375f), x);
    r = _mm256_fmadd_ps(n, _mm256_set1_ps(2.12194440e-4f), r);
    __m256 p = _mm256_set1_ps(1.9875691500e-4f);
    p = _mm256_fmadd_ps(p, r, _mm256_set1_ps(1.3981999507e-3f));
    p = _mm256_fmadd_ps(p, r, _mm256_set1_ps(8.3334519073e-3f));
    p = _mm256_fmadd_ps(p, r, _mm256_set1_ps(4.1665795894e-2f));
    p = _mm256_fmadd_ps(p, r, _mm256_set1_ps(1.6666665459e-1f));
    p = _mm256_fmadd_ps(p, r, _mm256_set1_ps(5.0000001201e-1f));
    const __m256 r2 = _mm256_mul_ps(r, r);
    const __m256 e = _mm256_add_ps(_mm256_fmadd_ps(p, r2, r), _mm256_set1_ps(1.0f));
    const __m256i ni = _mm256_cvttps_epi32(n);
    const __m256 res = _mm256_castsi256_ps(_mm256_add_epi32(_mm256_castps_si256(e), _mm256_slli_epi32(ni, 23)));
    return _mm256_andnot_ps(lt, res);
}
static inline __m256 orcb_gate8(__m256 f, __m256 g)
{
    const __m256 sign = _mm256_castsi256_ps(_mm256_set1_epi32((int)0x80000000u));
    const __m256 one = _mm256_set1_ps(1.0f);
    /* sigmoid32(g) = 1 / (1 + exp32(-g)) */
    const __m256 sg = _mm256_div_ps(one, _mm256_add_ps(one, orcb_exp32x8(_mm256_xor_ps(g, sign))));
    /* tanh32(f) = copysign(|f| > 44 ? 1 : 1 - 2 / (exp32(2|f|) + 1), f) */
    const __m256 af = _mm256_andnot_ps(sign, f);
    const __m256 e = orcb_exp32x8(_mm256_add_ps(af, af));
    __m256 r = _mm256_sub_ps(one, _mm256_div_ps(_mm256_set1_ps(2.0f), _mm256_add_ps(e, one)));
    r = _mm256_blendv_ps(r, one, _mm256_cmp_ps(af, _mm256_set1_ps(44.0f), _CMP_GT_OQ));
    const __m256 th = _mm256_or_ps(_mm256_andnot_ps(sign, r), _mm256_and_ps(sign, f));
    return _mm256_mul_ps(th, sg);
}
/* z[i] = tanh32(f[i]) * sigmoid32(g[i]), i < n (model.py:86) */
static void orcb_gate(const float *f, const float *g, float *z, int n)
{
    int i = 0;
    for (; i + 8 <= n; i += 8) _mm256_storeu_ps(z + i, orcb_gate8(_mm256_loadu_ps(f + i), _mm256_loadu_ps(g + i)));
    for (; i < n; ++i) z[i] = orc_tanh32(f[i]) * orc_sigmoid32(g[i]);
}
/* test hook: vector path (whole vectors + scalar tail) and the scalar functions on the same inputs */
void orc_gate_probe(const float *f, const float *g, int n, float *z_vec, float *z_scalar)
{
    orcb_gate(f, g, z_vec, n);
    for (int i = 0; i < n; ++i) z_scalar[i] = orc_tanh32(f[i]) * orc_sigmoid32(g[i]);
}

/* A thread's private, contiguous copy of the column slice [c0, c1) of a (K, stride) matrix: its share of the weights is then
 * a dense block it streams from its own L2 every step (a strided view of the shared matrix uses 64 of every 512 bytes it touches). */
static float *orcb_pack(const float *W, int stride, int K, int c0, int c1)
{
    const int nc = c1 - c0;
    float *p = (float *)malloc(sizeof(float) * (size_t)(K > 0 ? K : 1) * (nc > 0 ? nc : 1));
    for (int k = 0; k < K; ++k)
        for (int o = 0; o < nc; ++o) p[(size_t)k * nc + o] = W[(size_t)k * stride + c0 + o];
    return p;
}
/* mvb on a packed slice */
static void mvbp(const float *Wp, int c0, int c1, const float *X, int ldx, int N, int K, int t, float *out, int ldo, float *scratch)
{
    if (c1 > c0) mvb(Wp - c0, c1 - c0, c0, c1, X, ldx, N, K, t, out, ldo, scratch);
}
typedef struct { float *wf0, *wg0, *wf1, *wg1, *lcf, *lcg, *ws, *wd; } orcb_layer_pack;

typedef struct {
    orc_model *m; orc_plan p; int T, n_forced; const float *forced; const float *lc_up; int t_lc, lc_shift;
    const int32_t *gc_ids; const void *uniforms; float temperature; float *out_samples, *out_logits;
    int nt;
    orcb_barrier bar;
    /* shared activations, (N, width) row-major */
    float *x, *xn, *z, *acc, *c1, *lc_prev, *cq, *biasf, *biasg, *xin;
    int *id_prev, *id_cur;
    float **ring;      /* per layer: (d, N, R) */
} orcb_ctx;

typedef struct { orcb_ctx *g; int tid; } orcb_arg;

static void orcb_slice(int width, int nt, int tid, int *c0, int *c1)
{
    /* contiguous slices, multiples of 8 columns where the width allows (whole AVX2 vectors) */
    int unit = (width % (8 * nt) == 0) ? 8 : 1;
    int units = width / unit, per = (units + nt - 1) / nt;
    int a = tid * per, b = a + per;
    if (a > units) a = units;
    if (b > units) b = units;
    *c0 = a * unit; *c1 = b * unit;
}

static void *orcb_worker(void *arg_)
{
    orcb_arg *arg = (orcb_arg *)arg_;
    orcb_ctx *g = arg->g;
    const int tid = arg->tid, nt = g->nt;
    orc_model *m = g->m;
    const orc_config *c = &m->cfg;
    const orc_plan p = g->p;
    const int N = c->batch, L = c->n_layers, R = c->residual_channels, D = c->dilation_channels;
    const int S = c->skip_channels, C = c->lc_channels, O = m->out_dim, G = c->gc_channels;
    const int ifw = c->initial_filter_width, Q = c->quantization_channels, nr_mix = O / 3;
    const int T = g->T;
    int sense = 0;
    float *x = g->x, *xn = g->xn;          /* every thread swaps its own copy of the two pointers: no extra barrier */
    int d0, d1, r0, r1, s0, s1;
    orcb_slice(D, nt, tid, &d0, &d1);
    orcb_slice(R, nt, tid, &r0, &r1);
    orcb_slice(S, nt, tid, &s0, &s1);
    int maxc = S; if (D > maxc) maxc = D; if (R > maxc) maxc = R; if (O > maxc) maxc = O;
    float *scratch = (float *)malloc(sizeof(float) * 64 * (size_t)maxc * (N > 1 ? N : 1));   /* mvb: N * t * nc, t <= 64 */
    float *tmp = (float *)malloc(sizeof(float) * (size_t)N * maxc);
    float *f = (float *)malloc(sizeof(float) * (size_t)N * D), *gg = (float *)malloc(sizeof(float) * (size_t)N * D);
    float *c2 = (float *)malloc(sizeof(float) * O);
    float *gvec = (float *)calloc(G ? G : 1, sizeof(float));

    /* global conditioning folded into per-row biases (model.py:71-73), rows spread over threads */
    for (int b = tid; b < N; b += nt) {
        if (G) memcpy(gvec, m->gc_table + (size_t)g->gc_ids[b] * G, sizeof(float) * G);
        for (int l = 0; l < L; ++l) {
            float *bf = g->biasf + ((size_t)l * N + b) * D, *bg = g->biasg + ((size_t)l * N + b) * D;
            for (int o = 0; o < D; ++o) { bf[o] = m->layers[l].bf[o]; bg[o] = m->layers[l].bg[o]; }
            if (G) {
                mv_plan(m->layers[l].gcf, D, D, gvec, G, p.t_gc, tmp, scratch);
                for (int o = 0; o < D; ++o) bf[o] = bf[o] + tmp[o];
                mv_plan(m->layers[l].gcg, D, D, gvec, G, p.t_gc, tmp, scratch);
                for (int o = 0; o < D; ++o) bg[o] = bg[o] + tmp[o];
            }
        }
    }
    /* this thread's share of the weights, packed */
    orcb_layer_pack *pk = (orcb_layer_pack *)malloc(sizeof(orcb_layer_pack) * L);
    for (int l = 0; l < L; ++l) {
        const orc_layer *ly = &m->layers[l];
        pk[l].wf0 = orcb_pack(ly->wf, D, R, d0, d1);
        pk[l].wg0 = orcb_pack(ly->wg, D, R, d0, d1);
        pk[l].wf1 = orcb_pack(ly->wf + (size_t)R * D, D, R, d0, d1);
        pk[l].wg1 = orcb_pack(ly->wg + (size_t)R * D, D, R, d0, d1);
        pk[l].lcf = C ? orcb_pack(ly->lcf, D, C, d0, d1) : NULL;
        pk[l].lcg = C ? orcb_pack(ly->lcg, D, C, d0, d1) : NULL;
        pk[l].ws = orcb_pack(ly->ws, S, D, s0, s1);
        pk[l].wd = orcb_pack(ly->wd, R, D, r0, r1);
    }
    float *pk_w1 = orcb_pack(m->w1, S, S, s0, s1);
    float *pk_wc = c->scalar_input ? orcb_pack(m->wc, R, ifw, r0, r1) : NULL;
    orcb_wait(&g->bar, &sense);

#define ORCB_PUSH_INPUT(b)                                                 \
    do {                                                                   \
        if (c->scalar_input) {                                             \
            float *cq = g->cq + (size_t)(b) * ifw;                         \
            memmove(cq, cq + 1, sizeof(float) * (ifw - 1));                \
            cq[ifw - 1] = g->xin[b];                                       \
        } else {                                                           \
            g->id_prev[b] = g->id_cur[b];                                  \
            g->id_cur[b] = (int)g->xin[b];                                 \
        }                                                                  \
    } while (0)
    /* causal queue (model.py:122): per row, pushed by the row's owner -- here for step 0, after each draw for the next step */
    for (int b = tid; b < N; b += nt) ORCB_PUSH_INPUT(b);
    orcb_wait(&g->bar, &sense);

    for (int t = 0; t < T; ++t) {
        /* --- causal conv (model.py:131) by column slices over all rows --- */
        if (c->scalar_input) {
            mvbp(pk_wc, r0, r1, g->cq, ifw, N, ifw, p.t_causal, x, R, scratch);
        } else {
            for (int b = 0; b < N; ++b)
                for (int r = r0; r < r1; ++r) {
                    const int ip = g->id_prev[b], ic = g->id_cur[b];
                    float a = (ip >= 0) ? m->wc[((size_t)0 * Q + ip) * R + r] : 0.0f;
                    float bb = (ic >= 0 && ic < Q) ? m->wc[((size_t)1 * Q + ic) * R + r] : 0.0f;
                    x[(size_t)b * R + r] = a + bb;
                }
        }
        orcb_wait(&g->bar, &sense);
        /* --- dilated stack (model.py:141-149, 66-101) --- */
        for (int l = 0; l < L; ++l) {
            const orc_layer *ly = &m->layers[l];
            const int d = c->dilations[l];
            float *slot = g->ring[l] + (size_t)(t % d) * N * R;          /* (N, R): x_l(t - d) */
            const int nd = d1 - d0;
            if (nd > 0) {
                /* f = ((bias(+gc) + W0.old) (+ Wlc.lc)) + W1.cur, same for g (model.py:68-83), for this thread's channels */
                for (int b = 0; b < N; ++b)
                    for (int o = d0; o < d1; ++o) {
                        f[(size_t)b * D + o] = g->biasf[((size_t)l * N + b) * D + o];
                        gg[(size_t)b * D + o] = g->biasg[((size_t)l * N + b) * D + o];
                    }
#define ORCB_ADD(dst, Wm, Xm, ldx, K, tt)                                                      \
    do {                                                                                       \
        mvbp(Wm, d0, d1, Xm, ldx, N, K, tt, tmp, D, scratch);                                  \
        for (int b = 0; b < N; ++b)                                                            \
            for (int o = d0; o < d1; ++o) dst[(size_t)b * D + o] = dst[(size_t)b * D + o] + tmp[(size_t)b * D + o]; \
    } while (0)
                ORCB_ADD(f, pk[l].wf0, slot, R, R, p.t_old);
                ORCB_ADD(gg, pk[l].wg0, slot, R, R, p.t_old);
                if (C) {
                    ORCB_ADD(f, pk[l].lcf, g->lc_prev, C, C, p.t_lc);
                    ORCB_ADD(gg, pk[l].lcg, g->lc_prev, C, C, p.t_lc);
                }
                ORCB_ADD(f, pk[l].wf1, x, R, R, p.t_cur);
                ORCB_ADD(gg, pk[l].wg1, x, R, R, p.t_cur);
#undef ORCB_ADD
                for (int b = 0; b < N; ++b)                                                    /* model.py:86 */
                    orcb_gate(f + (size_t)b * D + d0, gg + (size_t)b * D + d0, g->z + (size_t)b * D + d0, nd);
            }
            orcb_wait(&g->bar, &sense);          /* z complete; every thread has read the ring slot and x */
            /* queue push (model.py:145), skip 1x1 + running sum (model.py:157), dense 1x1 + residual */
            for (int b = 0; b < N; ++b) memcpy(slot + (size_t)b * R + r0, x + (size_t)b * R + r0, sizeof(float) * (r1 - r0));
            if (s1 > s0) {
                mvbp(pk[l].ws, s0, s1, g->z, D, N, D, p.t_skip, tmp, S, scratch);
                for (int b = 0; b < N; ++b)
                    for (int s = s0; s < s1; ++s) {
                        float v = ly->bs[s] + tmp[(size_t)b * S + s];
                        float *a = g->acc + (size_t)b * S + s;
                        *a = (l == 0) ? v : *a + v;
                        if (l == L - 1) *a = relu32(*a);
                    }
            }
            if (r1 > r0) {
                const int dm = D / p.M;
                for (int b = 0; b < N; ++b)
                    for (int r = r0; r < r1; ++r) xn[(size_t)b * R + r] = x[(size_t)b * R + r] + ly->bd[r];
                for (int mm = 0; mm < p.M; ++mm) {
                    mvbp(pk[l].wd + (size_t)mm * dm * (r1 - r0), r0, r1, g->z + mm * dm, D, N, dm, p.t_dense, tmp, R, scratch);
                    for (int b = 0; b < N; ++b)
                        for (int r = r0; r < r1; ++r) xn[(size_t)b * R + r] = xn[(size_t)b * R + r] + tmp[(size_t)b * R + r];
                }
            }
            orcb_wait(&g->bar, &sense);          /* xn complete, every thread is done with x */
            { float *sw = x; x = xn; xn = sw; }
        }
        /* --- postprocessing (model.py:150-165): conv1 by column slices, conv2 + draw one row per thread --- */
        if (s1 > s0) {
            mvbp(pk_w1, s0, s1, g->acc, S, N, S, p.t_post1, tmp, S, scratch);
            for (int b = 0; b < N; ++b)
                for (int s = s0; s < s1; ++s) g->c1[(size_t)b * S + s] = relu32(m->b1[s] + tmp[(size_t)b * S + s]);
        }
        orcb_wait(&g->bar, &sense);
        for (int b = tid; b < N; b += nt) {
            const int sm = S / p.Mt;
            for (int o = 0; o < O; ++o) c2[o] = m->b2[o];
            for (int mm = 0; mm < p.Mt; ++mm) {
                mv_plan(m->w2 + (size_t)mm * sm * O, O, O, g->c1 + (size_t)b * S + mm * sm, sm, p.t_post2, tmp, scratch);
                for (int o = 0; o < O; ++o) c2[o] = c2[o] + tmp[o];
            }
            if (g->out_logits) memcpy(g->out_logits + ((size_t)b * T + t) * O, c2, sizeof(float) * O);
            float sample;
            if (c->scalar_input) {
                const float *u = (const float *)g->uniforms + ((size_t)b * T + t) * (nr_mix + 1);
                sample = mol_draw(c2, nr_mix, u);
            } else {
                double u = ((const double *)g->uniforms)[(size_t)b * T + t];
                sample = (float)mulaw_draw(c2, Q, g->temperature, u, NULL);
            }
            g->out_samples[(size_t)b * T + t] = sample;
            g->xin[b] = (t + 1 < g->n_forced) ? g->forced[(size_t)b * g->n_forced + t + 1] : sample;
            ORCB_PUSH_INPUT(b);
            /* lc queue push (model.py:125): the row for this step becomes lq[0] next step */
            if (C) {
                long idx = (long)t - g->lc_shift;
                float *dst = g->lc_prev + (size_t)b * C;
                if (g->lc_up && idx >= 0 && idx < g->t_lc) memcpy(dst, g->lc_up + ((size_t)b * g->t_lc + idx) * C, sizeof(float) * C);
                else memset(dst, 0, sizeof(float) * C);
            }
        }
        orcb_wait(&g->bar, &sense);
    }
#undef ORCB_PUSH_INPUT
    for (int l = 0; l < L; ++l) {
        free(pk[l].wf0); free(pk[l].wg0); free(pk[l].wf1); free(pk[l].wg1); free(pk[l].lcf); free(pk[l].lcg); free(pk[l].ws); free(pk[l].wd);
    }
    free(pk); free(pk_w1); free(pk_wc);
    free(scratch); free(tmp); free(f); free(gg); free(c2); free(gvec);
    return NULL;
}

/* Same contract as orc_generate (same plan -> bit-identical outputs); `threads` = worker threads (weight slices). */
int orc_generate_best(orc_model *m, const orc_plan *plan, int T, int n_forced, const float *forced,
                      const float *lc_up, int t_lc, int lc_shift, const int32_t *gc_ids,
                      const void *uniforms, float temperature, float *out_samples, float *out_logits, int threads)
{
    if (check_complete(m)) return -1;
    const orc_config *c = &m->cfg;
    const int N = c->batch, L = c->n_layers, R = c->residual_channels, D = c->dilation_channels;
    const int S = c->skip_channels, C = c->lc_channels, ifw = c->initial_filter_width, Q = c->quantization_channels;
    orc_plan p = *plan;
    if (!is_pow2(p.M) || D % p.M || !is_pow2(p.Mt) || S % p.Mt) { snprintf(m->err, sizeof m->err, "bad plan M/Mt"); return -1; }
    if (n_forced < 1) { snprintf(m->err, sizeof m->err, "n_forced must be >= 1"); return -1; }
    if (!c->scalar_input && !is_pow2(Q)) { snprintf(m->err, sizeof m->err, "Q must be a power of two"); return -1; }
    if (c->gc_channels && !gc_ids) { snprintf(m->err, sizeof m->err, "gc_ids required"); return -1; }
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;

    orcb_ctx g;
    memset(&g, 0, sizeof g);
    g.m = m; g.p = p; g.T = T; g.n_forced = n_forced; g.forced = forced; g.lc_up = lc_up; g.t_lc = t_lc; g.lc_shift = lc_shift;
    g.gc_ids = gc_ids; g.uniforms = uniforms; g.temperature = temperature; g.out_samples = out_samples; g.out_logits = out_logits;
    g.nt = threads;
    atomic_init(&g.bar.count, 0);
    atomic_init(&g.bar.sense, 0);
    g.bar.n = threads;
    g.x = zeros((long)N * R); g.xn = zeros((long)N * R); g.z = zeros((long)N * D); g.acc = zeros((long)N * S); g.c1 = zeros((long)N * S);
    g.lc_prev = zeros((long)N * (C ? C : 1)); g.cq = zeros((long)N * ifw);
    g.biasf = zeros((long)L * N * D); g.biasg = zeros((long)L * N * D); g.xin = zeros(N);
    g.id_prev = (int *)malloc(sizeof(int) * N); g.id_cur = (int *)malloc(sizeof(int) * N);
    g.ring = (float **)malloc(sizeof(float *) * L);
    for (int l = 0; l < L; ++l) g.ring[l] = zeros((long)c->dilations[l] * N * R);
    for (int b = 0; b < N; ++b) { g.id_prev[b] = -1; g.id_cur[b] = -1; g.xin[b] = forced[(size_t)b * n_forced]; }
    pthread_t th[256];
    orcb_arg args[256];
    for (int i = 0; i < threads; ++i) { args[i].g = &g; args[i].tid = i; }
    for (int i = 1; i < threads; ++i) pthread_create(&th[i], NULL, orcb_worker, &args[i]);
    orcb_worker(&args[0]);
    for (int i = 1; i < threads; ++i) pthread_join(th[i], NULL);
    for (int l = 0; l < L; ++l) free(g.ring[l]);
    free(g.ring); free(g.x); free(g.xn); free(g.z); free(g.acc); free(g.c1); free(g.lc_prev); free(g.cq); free(g.biasf); free(g.biasg);
    free(g.xin); free(g.id_prev); free(g.id_cur);
    return 0;
}
