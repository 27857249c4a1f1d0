/*
 * oracle/wn_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement (plain C) of the
 * reference's WaveNet incremental ("fast generation") sample loop.
 *
 * PARITY: pinned to the reference's OWN Python, not to TensorFlow's kernels.  The reference
 * (hccho2/Tacotron-Wavenet-Vocoder-Korean) ships no tests, golden vectors or checkpoints and
 * needs TensorFlow 1.x, which cannot be installed here.  Its wavenet/model.py, mixture.py and
 * ops.py are therefore imported unmodified from /root/reference and executed on a numpy
 * stand-in for the TF API (tests/golden/tf_numpy_shim.py); the outputs are committed as
 * tests/golden/ref_mol.npz / ref_mulaw.npz with the generating script
 * (tests/golden/make_reference_goldens.py) and this oracle reproduces them to 2e-5 on the
 * logits, samples and probabilities (tests/test_reference_pin.py), including every variable
 * name / shape and the queue layout.  The whole loop is pinned too: the reference's generate.py
 * main() was run unmodified on the stand-in (tests/golden/make_reference_generate_golden.py ->
 * ref_generate_main.npz) and orc_generate reproduces its waveforms from the same seeds -- mu-law
 * integer samples identical at temperature 1.0 and 0.7, the MoL waveform within 2.1e-7.
 * What stays UNPINNED is the arithmetic inside TF's own
 * kernels (conv1d, conv2d_transpose, softmax, random_uniform), restated from their published
 * definitions.  Further pins: (a) the known-answer values in the reference's comments,
 * (b) an independent numpy restatement (oracle/np_oracle.py), (c) torch's conv_transpose2d
 * for the upsampling network.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product never does.
 *
 * What is restated (reference file:line):
 *   wavenet/model.py:31-39    calculate_receptive_field      -> orc_receptive_field
 *   wavenet/model.py:41-46    _create_causal_layer           -> causal stage of orc_step
 *   wavenet/model.py:49-64    _create_queue / queue_initializer -> rings, zeroed per generate
 *   wavenet/model.py:66-101   _create_dilation_layer         -> layer stage of orc_step
 *   wavenet/model.py:102-111  create_upsample                -> orc_upsample
 *   wavenet/model.py:112-167  _create_network (train_mode=False) -> orc_step
 *   wavenet/model.py:181-212  _embed_gc                      -> gc table row
 *   wavenet/model.py:215-245  predict_proba_incremental      -> orc_step + head
 *   wavenet/mixture.py:84-114 sample_from_discretized_mix_logistic -> mol_draw
 *   generate.py:202-233       per-sample loop + categorical draw   -> orc_generate
 *   wavenet/ops.py:22-47      mu_law_encode / mu_law_decode  -> orc_mu_law_*
 *
 * Floating point evaluation order.  TF/Eigen's summation order inside conv1d is
 * not specified, so any order is an equally faithful restatement.  Every dot
 * product here goes through mv_plan(), whose order is selected by an orc_plan:
 * plan = all ones is the natural left-to-right order; a device kernel that wants
 * bit-exact comparison reports the plan it implements and the test passes it in.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include "wn_math_ref.h"

#define ORC_MAX_LAYERS 256
#define ORC_MAX_UP 8

typedef struct {
    int32_t batch;
    int32_t n_layers;
    int32_t filter_width;            /* only 2 is supported (hparams.py:59) */
    int32_t residual_channels;       /* R */
    int32_t dilation_channels;       /* D */
    int32_t skip_channels;           /* S */
    int32_t quantization_channels;   /* Q */
    int32_t out_channels;            /* O (scalar_input) */
    int32_t use_biases;
    int32_t scalar_input;
    int32_t initial_filter_width;    /* ifw */
    int32_t gc_channels;             /* G, 0 = no global conditioning */
    int32_t gc_cardinality;
    int32_t lc_channels;             /* C, 0 = no local conditioning */
    int32_t n_upsample;
    int32_t upsample_factor[ORC_MAX_UP];
    int32_t dilations[ORC_MAX_LAYERS];
} orc_config;

/* Evaluation-order plan, see mv_plan(). */
typedef struct {
    int32_t M;         /* dense (residual 1x1): K=D cut in M slices added in order   */
    int32_t Mt;        /* post2: K=S cut in Mt slices added in order                 */
    int32_t t_cur;     /* filter/gate, current tap (K=R)                             */
    int32_t t_old;     /* filter/gate, dilated tap (K=R)                             */
    int32_t t_lc;      /* lc_filter/lc_gate (K=C)                                    */
    int32_t t_gc;      /* gc_filter/gc_gate (K=G)                                    */
    int32_t t_dense;   /* per slice of the dense 1x1 (K=D/M)                         */
    int32_t t_skip;    /* skip 1x1 (K=D)                                             */
    int32_t t_post1;   /* postprocess conv1 (K=S)                                    */
    int32_t t_post2;   /* per slice of postprocess conv2 (K=S/Mt)                    */
    int32_t t_causal;  /* scalar causal conv (K=ifw)                                 */
} orc_plan;

typedef struct {
    float *wf, *wg;      /* (2, R, D) conv_filter / conv_gate kernels */
    float *bf, *bg;      /* (D) */
    float *gcf, *gcg;    /* (G, D) */
    float *lcf, *lcg;    /* (C, D) */
    float *wd, *bd;      /* (D, R), (R) */
    float *ws, *bs;      /* (D, S), (S) */
} orc_layer;

typedef struct {
    orc_config cfg;
    int out_dim;                 /* O or Q */
    float *gc_table;             /* (card, G) */
    float *up[ORC_MAX_UP];       /* (F, 2) each */
    float *wc;                   /* causal kernel (ifw,1,R) or (2,Q,R) */
    orc_layer *layers;
    float *w1, *b1;              /* (S,S),(S) */
    float *w2, *b2;              /* (S,out),(out) */
    char err[256];
} orc_model;

/* ------------------------------------------------------------------------- */
/* mv_plan: out[o] = sum_k W[k*stride + o] * x[k] for o < ncols, k < K.  Every column is evaluated
 * as t contiguous chunks of K/t; each chunk is an fma chain in increasing k starting from +0; the t
 * chunk sums are combined by an xor-butterfly with ascending offsets (1,2,4,...,t/2):
 * a[c] = a[c] + a[c^off].  t must be a power of two dividing K.  The loops run k-outer / column-inner
 * so the compiler can vectorise across columns; each column's operation sequence is unchanged.
 * scratch: t*ncols floats. */
static void mv_plan(const float *W, int stride, int ncols, const float *x, int K, int t, float *out, float *scratch)
{
    int ch = K / t;
    for (int c = 0; c < t; ++c) {
        float *restrict a = scratch + (size_t)c * ncols;
        for (int o = 0; o < ncols; ++o) a[o] = 0.0f;
        for (int i = 0; i < ch; ++i) {
            int k = c * ch + i;
            const float xk = x[k];
            const float *restrict w = W + (size_t)k * stride;
            for (int o = 0; o < ncols; ++o) a[o] = fmaf(w[o], xk, a[o]);
        }
    }
    for (int off = 1; off < t; off <<= 1)
        for (int c = 0; c < t; ++c)
            if ((c & off) == 0) {
                float *restrict a = scratch + (size_t)c * ncols;
                float *restrict b = scratch + (size_t)(c ^ off) * ncols;
                for (int o = 0; o < ncols; ++o) { float v = a[o] + b[o]; a[o] = v; b[o] = v; }
            }
    for (int o = 0; o < ncols; ++o) out[o] = scratch[o];
}

/* ascending xor-butterfly sum of n (power of two) doubles; used for the softmax
 * denominator (model.py:243). */
static double tree_sum64(const double *v, int n)
{
    double *a = (double *)malloc(sizeof(double) * n * 2);
    double *b = a + n;
    memcpy(a, v, sizeof(double) * n);
    for (int off = 1; off < n; off <<= 1) {
        for (int c = 0; c < n; ++c) b[c] = a[c] + a[c ^ off];
        memcpy(a, b, sizeof(double) * n);
    }
    double r = a[0];
    free(a);
    return r;
}

/* ascending xor-butterfly logaddexp reduction (generate.py:221; numpy reduces
 * left to right -- the tree order is the pinned, parallel-friendly variant). */
static float tree_logaddexp32(const float *v, int n)
{
    float *a = (float *)malloc(sizeof(float) * n * 2);
    float *b = a + n;
    memcpy(a, v, sizeof(float) * n);
    for (int off = 1; off < n; off <<= 1) {
        for (int c = 0; c < n; ++c) {
            int o = c ^ off;
            /* evaluate with the lower index first so both partners agree bitwise */
            b[c] = (c < o) ? orc_logaddexp32(a[c], a[o]) : orc_logaddexp32(a[o], a[c]);
        }
        memcpy(a, b, sizeof(float) * n);
    }
    float r = a[0];
    free(a);
    return r;
}

static int is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

/* ------------------------------------------------------------------------- */
int orc_receptive_field(int filter_width, const int32_t *dilations, int n, int scalar_input,
                        int initial_filter_width)
{
    /* wavenet/model.py:31-39 */
    int s = 0;
    for (int i = 0; i < n; ++i) s += dilations[i];
    int rf = (filter_width - 1) * s + 1;
    rf += scalar_input ? initial_filter_width - 1 : filter_width - 1;
    return rf;
}

void orc_mu_law_encode(const float *audio, int n, int quantization_channels, int32_t *out)
{
    /* wavenet/ops.py:22-33; TF evaluates in fp32 with its own log1p; tf.to_int32 truncates.  The output is an integer
     * code, so the evaluation is pinned (orc_log1p32, wn_math_ref.h) and the CUDA kernel must reproduce it bit for bit;
     * against the reference's own run (libm log1p) the code can differ by 1 only where the argument sits on a cell edge. */
    float mu = (float)(quantization_channels - 1);
    float den = orc_log1p32(mu);
    for (int i = 0; i < n; ++i) {
        float a = audio[i];
        float safe = fminf(fabsf(a), 1.0f);
        float mag = orc_log1p32(mu * safe) / den;
        float sgn = (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : 0.0f);
        float sig = sgn * mag;
        out[i] = (int32_t)((sig + 1.0f) / 2.0f * mu + 0.5f);
    }
}

void orc_mu_law_decode(const float *in, int n, int quantization_channels, int quantization, float *out)
{
    /* wavenet/ops.py:36-47; `in` holds ids (as floats) when quantization!=0, else
     * companded values in [-1,1]. */
    float mu = (float)(quantization_channels - 1);
    for (int i = 0; i < n; ++i) {
        float sig = quantization ? 2.0f * (in[i] / mu) - 1.0f : in[i];
        float mag = (1.0f / mu) * (powf(1.0f + mu, fabsf(sig)) - 1.0f);
        float sgn = (sig > 0.0f) ? 1.0f : ((sig < 0.0f) ? -1.0f : 0.0f);
        out[i] = sgn * mag;
    }
}

/* ------------------------------------------------------------------------- */
orc_model *orc_create(const orc_config *cfg)
{
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->cfg = *cfg;
    m->out_dim = cfg->scalar_input ? cfg->out_channels : cfg->quantization_channels;
    m->layers = (orc_layer *)calloc(cfg->n_layers, sizeof(orc_layer));
    return m;
}

static void free_layer(orc_layer *l)
{
    free(l->wf); free(l->wg); free(l->bf); free(l->bg); free(l->gcf); free(l->gcg);
    free(l->lcf); free(l->lcg); free(l->wd); free(l->bd); free(l->ws); free(l->bs);
}

void orc_destroy(orc_model *m)
{
    if (!m) return;
    for (int i = 0; i < m->cfg.n_layers; ++i) free_layer(&m->layers[i]);
    free(m->layers);
    free(m->gc_table); free(m->wc); free(m->w1); free(m->b1); free(m->w2); free(m->b2);
    for (int i = 0; i < ORC_MAX_UP; ++i) free(m->up[i]);
    free(m);
}

const char *orc_last_error(orc_model *m) { return m->err; }

static int set_buf(orc_model *m, float **dst, const float *src, long n, long expect, const char *name)
{
    if (n != expect) {
        snprintf(m->err, sizeof m->err, "weight %s: got %ld elements, expected %ld", name, n, expect);
        return -1;
    }
    free(*dst);
    *dst = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(*dst, src, sizeof(float) * (size_t)n);
    return 0;
}

/* Names are the TF variable names of the reference graph (scope "wavenet/",
 * model.py:221; tf.layers auto-numbering), see SURVEY.md Appendix B. */
int orc_set_weight(orc_model *m, const char *name, const float *data, long n)
{
    const orc_config *c = &m->cfg;
    long R = c->residual_channels, D = c->dilation_channels, S = c->skip_channels;
    long G = c->gc_channels, C = c->lc_channels;
    int li; char tail[128];
    if (!strcmp(name, "wavenet/gc_embedding"))
        return set_buf(m, &m->gc_table, data, n, (long)c->gc_cardinality * G, name);
    if (sscanf(name, "wavenet/upsample%d/kernel", &li) == 1 && li >= 0 && li < c->n_upsample)
        return set_buf(m, &m->up[li], data, n, (long)c->upsample_factor[li] * 2, name);
    if (!strcmp(name, "wavenet/conv1d/kernel"))
        return set_buf(m, &m->wc, data, n,
                       c->scalar_input ? (long)c->initial_filter_width * R : 2L * c->quantization_channels * R, name);
    if (!strcmp(name, "wavenet/conv1d_1/kernel")) return set_buf(m, &m->w1, data, n, S * S, name);
    if (!strcmp(name, "wavenet/conv1d_1/bias")) return set_buf(m, &m->b1, data, n, S, name);
    if (!strcmp(name, "wavenet/conv1d_2/kernel")) return set_buf(m, &m->w2, data, n, S * m->out_dim, name);
    if (!strcmp(name, "wavenet/conv1d_2/bias")) return set_buf(m, &m->b2, data, n, m->out_dim, name);
    if (sscanf(name, "wavenet/dilated_stack/layer%d/dilation_layer/%127s", &li, tail) == 2 &&
        li >= 0 && li < c->n_layers) {
        orc_layer *l = &m->layers[li];
        if (!strcmp(tail, "conv_filter/kernel")) return set_buf(m, &l->wf, data, n, 2 * R * D, name);
        if (!strcmp(tail, "conv_filter/bias")) return set_buf(m, &l->bf, data, n, D, name);
        if (!strcmp(tail, "conv_gate/kernel")) return set_buf(m, &l->wg, data, n, 2 * R * D, name);
        if (!strcmp(tail, "conv_gate/bias")) return set_buf(m, &l->bg, data, n, D, name);
        if (!strcmp(tail, "gc_filter/kernel")) return set_buf(m, &l->gcf, data, n, G * D, name);
        if (!strcmp(tail, "gc_gate/kernel")) return set_buf(m, &l->gcg, data, n, G * D, name);
        if (!strcmp(tail, "lc_filter/kernel")) return set_buf(m, &l->lcf, data, n, C * D, name);
        if (!strcmp(tail, "lc_gate/kernel")) return set_buf(m, &l->lcg, data, n, C * D, name);
        if (!strcmp(tail, "dense/kernel")) return set_buf(m, &l->wd, data, n, D * R, name);
        if (!strcmp(tail, "dense/bias")) return set_buf(m, &l->bd, data, n, R, name);
        if (!strcmp(tail, "skip/kernel")) return set_buf(m, &l->ws, data, n, D * S, name);
        if (!strcmp(tail, "skip/bias")) return set_buf(m, &l->bs, data, n, S, name);
    }
    snprintf(m->err, sizeof m->err, "unknown weight name %s", name);
    return -1;
}

static float *zeros(long n) { return (float *)calloc((size_t)n, sizeof(float)); }

static int check_complete(orc_model *m)
{
    const orc_config *c = &m->cfg;
    if (c->filter_width != 2) { snprintf(m->err, sizeof m->err, "filter_width must be 2"); return -1; }
    if (!m->wc || !m->w1 || !m->w2) { snprintf(m->err, sizeof m->err, "missing causal/post kernels"); return -1; }
    if (!m->b1) m->b1 = zeros(c->skip_channels);
    if (!m->b2) m->b2 = zeros(m->out_dim);
    if (c->gc_channels && !m->gc_table) { snprintf(m->err, sizeof m->err, "missing gc_embedding"); return -1; }
    for (int i = 0; i < c->n_layers; ++i) {
        orc_layer *l = &m->layers[i];
        if (!l->wf || !l->wg || !l->wd || !l->ws) { snprintf(m->err, sizeof m->err, "layer %d incomplete", i); return -1; }
        if (c->gc_channels && (!l->gcf || !l->gcg)) { snprintf(m->err, sizeof m->err, "layer %d gc missing", i); return -1; }
        if (c->lc_channels && (!l->lcf || !l->lcg)) { snprintf(m->err, sizeof m->err, "layer %d lc missing", i); return -1; }
        if (!l->bf) l->bf = zeros(c->dilation_channels);
        if (!l->bg) l->bg = zeros(c->dilation_channels);
        if (!l->bd) l->bd = zeros(c->residual_channels);
        if (!l->bs) l->bs = zeros(c->skip_channels);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* create_upsample, wavenet/model.py:102-111.  Each stage is
 * conv2d_transpose(filters=1, kernel=(F,2), strides=(F,1), 'same', no bias):
 *   out[i*F + a][w] = in[i][w]*K[a][0] + in[i][w-1]*K[a][1],  in[.][-1] = 0
 * (SURVEY.md A.3).  Pinned evaluation: v = in[i][w]*K[a][0]; v = fma(in[i][w-1], K[a][1], v).
 * mel: (T_mel, C) one batch row.  out: (T_mel*prod(F), C).                       */
int orc_upsample(orc_model *m, const float *mel, int t_mel, float *out)
{
    const orc_config *c = &m->cfg;
    int C = c->lc_channels;
    long T = t_mel;
    float *cur = (float *)malloc(sizeof(float) * (size_t)T * C);
    memcpy(cur, mel, sizeof(float) * (size_t)T * C);
    for (int s = 0; s < c->n_upsample; ++s) {
        int F = c->upsample_factor[s];
        const float *K = m->up[s];
        if (!K) { free(cur); snprintf(m->err, sizeof m->err, "missing upsample%d kernel", s); return -1; }
        float *nxt = (float *)malloc(sizeof(float) * (size_t)T * F * C);
        for (long i = 0; i < T; ++i)
            for (int a = 0; a < F; ++a)
                for (int w = 0; w < C; ++w) {
                    float v = cur[i * C + w] * K[a * 2 + 0];
                    float prev = (w > 0) ? cur[i * C + w - 1] : 0.0f;
                    v = fmaf(prev, K[a * 2 + 1], v);
                    nxt[(i * F + a) * C + w] = v;
                }
        free(cur);
        cur = nxt;
        T *= F;
    }
    memcpy(out, cur, sizeof(float) * (size_t)T * C);
    free(cur);
    return 0;
}

/* ------------------------------------------------------------------------- */
typedef struct {
    /* per batch row state; queues are rings (arithmetic identical to the
     * shift-copy queues of model.py:122,125,145) */
    float *cq;        /* scalar: last ifw inputs, oldest first */
    int id_prev, id_cur;
    float *lc_prev;   /* lq[0]: the previous step's lc row (model.py:79-80) */
    float **ring;     /* per layer: d rows of R */
    float *biasf, *biasg;  /* (L, D): bias + gc contribution, folded once */
} row_state;

static float relu32(float v) { return v > 0.0f ? v : 0.0f; }

/* mixture.py:84-114.  y: 3*nr_mix logits; u: nr_mix + 1 uniforms. */
static float mol_draw(const float *y, int nr_mix, const float *u)
{
    int best = 0; float bestv = 0.0f;
    for (int k = 0; k < nr_mix; ++k) {
        float g = y[k] - orc_log32(-orc_log32(u[k]));
        if (k == 0 || g > bestv) { bestv = g; best = k; }
    }
    float mean = y[nr_mix + best];
    float ls = y[2 * nr_mix + best];
    const float log_scale_min = -32.23619130191664f;   /* float(np.log(1e-14)) */
    if (!(ls > log_scale_min)) ls = log_scale_min;       /* tf.maximum */
    float u2 = u[nr_mix];
    float d = orc_log32(u2) - orc_log32(1.0f - u2);
    float x = mean + orc_exp32(ls) * d;
    x = fmaxf(x, -1.0f);
    x = fminf(x, 1.0f);
    return x;
}

/* model.py:243 + generate.py:219-231: float64 softmax -> fp32, temperature, draw */
static int mulaw_draw(const float *c2, int Q, float temperature, double u, float *probs_out)
{
    float mx = c2[0];
    for (int j = 1; j < Q; ++j) if (c2[j] > mx) mx = c2[j];
    double *e = (double *)malloc(sizeof(double) * Q);
    for (int j = 0; j < Q; ++j) e[j] = orc_exp64((double)c2[j] - (double)mx);
    double den = tree_sum64(e, Q);
    float *s = (float *)malloc(sizeof(float) * Q);
    for (int j = 0; j < Q; ++j) {
        float p = (float)(e[j] / den);
        if (probs_out) probs_out[j] = p;
        s[j] = orc_log32(p) / temperature;
    }
    float lse = tree_logaddexp32(s, Q);
    for (int j = 0; j < Q; ++j) e[j] = (double)orc_exp32(s[j] - lse);
    /* np.cumsum(float64(q)) in a pinned, parallel-friendly order: a Kogge-Stone scan inside each block
     * of 32 (v[j] += v[j-off], off = 1,2,..,16), then each block adds the left-to-right sum of the
     * preceding block totals.  numpy accumulates strictly left to right; the two differ by at most an
     * ulp of fp64, which moves a draw only when u lies within ~1e-16 of a bin edge. */
    {
        double tot[64];
        int nb = (Q + 31) / 32;
        for (int bk = 0; bk < nb; ++bk) {
            int blk = bk * 32;
            int n = (Q - blk < 32) ? (Q - blk) : 32;
            for (int off = 1; off < n; off <<= 1)
                for (int j = n - 1; j >= off; --j) e[blk + j] = e[blk + j] + e[blk + j - off];
            tot[bk] = e[blk + n - 1];
        }
        double base = tot[0];
        for (int bk = 1; bk < nb; ++bk) {
            int blk = bk * 32;
            int n = (Q - blk < 32) ? (Q - blk) : 32;
            for (int j = 0; j < n; ++j) e[blk + j] = base + e[blk + j];
            base = base + tot[bk];
        }
    }
    double acc = e[Q - 1];
    int cnt = 0;
    for (int j = 0; j < Q; ++j) if (e[j] / acc <= u) ++cnt;   /* searchsorted(side='right') */
    if (cnt > Q - 1) cnt = Q - 1;
    free(e); free(s);
    return cnt;
}

typedef struct {
    orc_model *m; orc_plan p; int T, n_forced; const float *forced; const float *lc_up; int t_lc, lc_shift;
    const int32_t *gc_ids; const void *uniforms; float temperature; float *out_samples, *out_logits;
    int next_row; pthread_mutex_t mu;
} gen_ctx;

static void run_row(gen_ctx *g_, int b)
{
    orc_model *m = g_->m;
    const orc_config *c = &m->cfg;
    const orc_plan p = g_->p;
    const int L = c->n_layers, R = c->residual_channels, D = c->dilation_channels;
    const int S = c->skip_channels, G = c->gc_channels, C = c->lc_channels, O = m->out_dim;
    const int ifw = c->initial_filter_width, Q = c->quantization_channels;
    const int T = g_->T, n_forced = g_->n_forced, t_lc = g_->t_lc, lc_shift = g_->lc_shift;
    const float *forced = g_->forced, *lc_up = g_->lc_up;
    const int32_t *gc_ids = g_->gc_ids;
    const void *uniforms = g_->uniforms;
    const float temperature = g_->temperature;
    float *out_samples = g_->out_samples, *out_logits = g_->out_logits;
    const int nr_mix = O / 3;
    int maxc = S; if (D > maxc) maxc = D; if (R > maxc) maxc = R; if (O > maxc) maxc = O;
    {
        float *x = (float *)malloc(sizeof(float) * R), *xn = (float *)malloc(sizeof(float) * R);
        float *z = (float *)malloc(sizeof(float) * D);
        float *f = (float *)malloc(sizeof(float) * D), *g = (float *)malloc(sizeof(float) * D);
        float *acc = (float *)malloc(sizeof(float) * S), *c1 = (float *)malloc(sizeof(float) * S);
        float *c2 = (float *)malloc(sizeof(float) * O);
        float *tmp = (float *)malloc(sizeof(float) * maxc);
        float *scratch = (float *)malloc(sizeof(float) * 64 * (size_t)maxc);
        float *gvec = (float *)calloc(G ? G : 1, sizeof(float));
        float *zero_lc = (float *)calloc(C ? C : 1, sizeof(float));
        row_state st;
        st.cq = (float *)calloc(ifw, sizeof(float));
        st.id_prev = -1; st.id_cur = -1;        /* zero one-hot rows (queue_initializer) */
        st.lc_prev = (float *)calloc(C ? C : 1, sizeof(float));
        st.ring = (float **)malloc(sizeof(float *) * L);
        for (int l = 0; l < L; ++l) st.ring[l] = (float *)calloc((size_t)c->dilations[l] * R, sizeof(float));
        st.biasf = (float *)malloc(sizeof(float) * L * D);
        st.biasg = (float *)malloc(sizeof(float) * L * D);
        if (G) memcpy(gvec, m->gc_table + (size_t)gc_ids[b] * G, sizeof(float) * G);
        for (int l = 0; l < L; ++l) {
            /* global conditioning is constant per row: fold it into the biases once (model.py:71-73) */
            for (int o = 0; o < D; ++o) { st.biasf[l * D + o] = m->layers[l].bf[o]; st.biasg[l * D + o] = m->layers[l].bg[o]; }
            if (G) {
                mv_plan(m->layers[l].gcf, D, D, gvec, G, p.t_gc, tmp, scratch);
                for (int o = 0; o < D; ++o) st.biasf[l * D + o] = st.biasf[l * D + o] + tmp[o];
                mv_plan(m->layers[l].gcg, D, D, gvec, G, p.t_gc, tmp, scratch);
                for (int o = 0; o < D; ++o) st.biasg[l * D + o] = st.biasg[l * D + o] + tmp[o];
            }
        }

        float prev_sample = 0.0f;
        for (int t = 0; t < T; ++t) {
            float x_in = (t < n_forced) ? forced[(size_t)b * n_forced + t] : prev_sample;
            /* --- causal queue + causal conv (model.py:122,131) --- */
            if (c->scalar_input) {
                memmove(st.cq, st.cq + 1, sizeof(float) * (ifw - 1));
                st.cq[ifw - 1] = x_in;
                mv_plan(m->wc, R, R, st.cq, ifw, p.t_causal, x, scratch);
            } else {
                st.id_prev = st.id_cur;
                st.id_cur = (int)x_in;
                for (int r = 0; r < R; ++r) {
                    float a = (st.id_prev >= 0) ? m->wc[((size_t)0 * Q + st.id_prev) * R + r] : 0.0f;
                    float bb = (st.id_cur >= 0 && st.id_cur < Q) ? m->wc[((size_t)1 * Q + st.id_cur) * R + r] : 0.0f;
                    x[r] = a + bb;
                }
            }
            const float *lc_use = st.lc_prev;     /* lq[0] */
            /* --- dilated stack (model.py:141-149, 66-101) --- */
            for (int l = 0; l < L; ++l) {
                const orc_layer *ly = &m->layers[l];
                int d = c->dilations[l];
                float *slot = st.ring[l] + (size_t)(t % d) * R;   /* holds x_l(t-d) */
                /* f = ((bias(+gc) + W0.old) (+ Wlc.lc)) + W1.cur, same for g (model.py:68-83) */
                memcpy(f, st.biasf + l * D, sizeof(float) * D);
                memcpy(g, st.biasg + l * D, sizeof(float) * D);
                mv_plan(ly->wf, D, D, slot, R, p.t_old, tmp, scratch);
                for (int o = 0; o < D; ++o) f[o] = f[o] + tmp[o];
                mv_plan(ly->wg, D, D, slot, R, p.t_old, tmp, scratch);
                for (int o = 0; o < D; ++o) g[o] = g[o] + tmp[o];
                if (C) {
                    mv_plan(ly->lcf, D, D, lc_use, C, p.t_lc, tmp, scratch);
                    for (int o = 0; o < D; ++o) f[o] = f[o] + tmp[o];
                    mv_plan(ly->lcg, D, D, lc_use, C, p.t_lc, tmp, scratch);
                    for (int o = 0; o < D; ++o) g[o] = g[o] + tmp[o];
                }
                mv_plan(ly->wf + (size_t)R * D, D, D, x, R, p.t_cur, tmp, scratch);
                for (int o = 0; o < D; ++o) f[o] = f[o] + tmp[o];
                mv_plan(ly->wg + (size_t)R * D, D, D, x, R, p.t_cur, tmp, scratch);
                for (int o = 0; o < D; ++o) g[o] = g[o] + tmp[o];
                for (int o = 0; o < D; ++o) z[o] = orc_tanh32(f[o]) * orc_sigmoid32(g[o]);   /* model.py:86 */
                memcpy(slot, x, sizeof(float) * R);                /* queue push, model.py:145 */
                mv_plan(ly->ws, S, S, z, D, p.t_skip, tmp, scratch);
                for (int s = 0; s < S; ++s) {
                    float v = ly->bs[s] + tmp[s];
                    acc[s] = (l == 0) ? v : acc[s] + v;             /* sum(outputs), model.py:157 */
                }
                int dm = D / p.M;
                for (int r = 0; r < R; ++r) xn[r] = x[r] + ly->bd[r];
                for (int mm = 0; mm < p.M; ++mm) {
                    mv_plan(ly->wd + (size_t)mm * dm * R, R, R, z + mm * dm, dm, p.t_dense, tmp, scratch);
                    for (int r = 0; r < R; ++r) xn[r] = xn[r] + tmp[r];
                }
                memcpy(x, xn, sizeof(float) * R);
            }
            /* --- postprocessing (model.py:150-165) --- */
            for (int s = 0; s < S; ++s) acc[s] = relu32(acc[s]);
            mv_plan(m->w1, S, S, acc, S, p.t_post1, tmp, scratch);
            for (int s = 0; s < S; ++s) c1[s] = relu32(m->b1[s] + tmp[s]);
            int sm = S / p.Mt;
            for (int o = 0; o < O; ++o) c2[o] = m->b2[o];
            for (int mm = 0; mm < p.Mt; ++mm) {
                mv_plan(m->w2 + (size_t)mm * sm * O, O, O, c1 + mm * sm, sm, p.t_post2, tmp, scratch);
                for (int o = 0; o < O; ++o) c2[o] = c2[o] + tmp[o];
            }
            if (out_logits) memcpy(out_logits + ((size_t)b * T + t) * O, c2, sizeof(float) * O);
            /* --- head + draw --- */
            float sample;
            if (c->scalar_input) {
                const float *u = (const float *)uniforms + ((size_t)b * T + t) * (nr_mix + 1);
                sample = mol_draw(c2, nr_mix, u);
            } else {
                double u = ((const double *)uniforms)[(size_t)b * T + t];
                sample = (float)mulaw_draw(c2, Q, temperature, u, NULL);
            }
            out_samples[(size_t)b * T + t] = sample;
            prev_sample = sample;
            /* --- lc queue push (model.py:125): the row for this step becomes lq[0] next step --- */
            if (C) {
                long idx = (long)t - lc_shift;
                const float *row = (lc_up && idx >= 0 && idx < t_lc) ? lc_up + ((size_t)b * t_lc + idx) * C : zero_lc;
                memcpy(st.lc_prev, row, sizeof(float) * C);
            }
        }
        for (int l = 0; l < L; ++l) free(st.ring[l]);
        free(st.ring); free(st.cq); free(st.lc_prev); free(st.biasf); free(st.biasg);
        free(x); free(xn); free(z); free(f); free(g); free(acc); free(c1); free(c2); free(tmp); free(scratch);
        free(gvec); free(zero_lc);
    }
}

static void *row_worker(void *arg)
{
    gen_ctx *g_ = (gen_ctx *)arg;
    for (;;) {
        pthread_mutex_lock(&g_->mu);
        int b = g_->next_row++;
        pthread_mutex_unlock(&g_->mu);
        if (b >= g_->m->cfg.batch) break;
        run_row(g_, b);
    }
    return NULL;
}

static int g_threads = 1;
/* number of host threads orc_generate spreads the (independent) batch rows over */
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* One generate.py-style run for the whole batch.
 *   T            number of network steps per row
 *   n_forced     x_in(t) = forced[b][t] for t < n_forced, else the sample drawn at t-1.
 *                n_forced >= 1 (forced[b][0] is the initial sample, generate.py:184-192);
 *                n_forced == T is teacher forcing; priming with a seed uses
 *                n_forced = len(seed) (generate.py:177-180).
 *   lc_up        (batch, T_lc, C) upsampled local condition or NULL
 *   lc_shift     step t pushes row lc_up[t - lc_shift] (zeros when negative, generate.py:180)
 *                and layers read the row pushed at step t-1 (model.py:79-80)
 *   gc_ids       (batch) or NULL
 *   uniforms     scalar: (batch, T, nr_mix+1) fp32;  mu-law: (batch, T) fp64
 *   out_samples  (batch, T) fp32 (mu-law ids stored as floats, as generate.py does)
 *   out_logits   optional (batch, T, out_dim) raw conv2 output
 */
int orc_generate(orc_model *m, const orc_plan *plan, int T, int n_forced, const float *forced,
                 const float *lc_up, int t_lc, int lc_shift, const int32_t *gc_ids,
                 const void *uniforms, float temperature, float *out_samples, float *out_logits)
{
    if (check_complete(m)) return -1;
    const orc_config *c = &m->cfg;
    const int N = c->batch, L = c->n_layers, R = c->residual_channels, D = c->dilation_channels;
    const int S = c->skip_channels, G = c->gc_channels, C = c->lc_channels;
    const int ifw = c->initial_filter_width, Q = c->quantization_channels;
    orc_plan p = *plan;
    if (!is_pow2(p.M) || D % p.M || !is_pow2(p.Mt) || S % p.Mt) { snprintf(m->err, sizeof m->err, "bad plan M/Mt"); return -1; }
    if (n_forced < 1) { snprintf(m->err, sizeof m->err, "n_forced must be >= 1"); return -1; }
    if (!c->scalar_input && !is_pow2(Q)) { snprintf(m->err, sizeof m->err, "Q must be a power of two"); return -1; }

    gen_ctx g_;
    g_.m = m; g_.p = p; g_.T = T; g_.n_forced = n_forced; g_.forced = forced; g_.lc_up = lc_up; g_.t_lc = t_lc;
    g_.lc_shift = lc_shift; g_.gc_ids = gc_ids; g_.uniforms = uniforms; g_.temperature = temperature;
    g_.out_samples = out_samples; g_.out_logits = out_logits; g_.next_row = 0;
    pthread_mutex_init(&g_.mu, NULL);
    int nt = g_threads < N ? g_threads : N;
    if (nt <= 1) {
        for (int b = 0; b < N; ++b) run_row(&g_, b);
    } else {
        pthread_t th[256];
        if (nt > 256) nt = 256;
        for (int i = 0; i < nt; ++i) pthread_create(&th[i], NULL, row_worker, &g_);
        for (int i = 0; i < nt; ++i) pthread_join(th[i], NULL);
    }
    pthread_mutex_destroy(&g_.mu);
    (void)L; (void)R; (void)G; (void)C; (void)ifw;
    return 0;
}

/* softmax probabilities of one logits row, exactly as the head computes them
 * (model.py:243); exposed for tests of the invariant at generate.py:227-228. */
void orc_softmax_probs(const float *c2, int Q, float *probs)
{
    mulaw_draw(c2, Q, 1.0f, 0.5, probs);
}

/* mixture.py:84-114 on a tensor of logits: y (rows, 3*nr_mix), u (rows, nr_mix + 1) -> out (rows); the same mol_draw the
 * generation loop calls, exposed for the test of the product's wn_mol_sample. */
void orc_mol_sample(const float *y, const float *u, long rows, int nr_mix, float *out)
{
    for (long i = 0; i < rows; ++i) out[i] = mol_draw(y + i * 3 * nr_mix, nr_mix, u + i * (nr_mix + 1));
}

float orc_math_probe(int which, float x)
{
    switch (which) {
    case 0: return orc_exp32(x);
    case 1: return orc_log32(x);
    case 2: return orc_tanh32(x);
    case 3: return orc_sigmoid32(x);
    case 4: return orc_log1p32(x);
    default: return (float)orc_exp64((double)x);
    }
}

/* the "best-effort CPU" baseline (BASELINE.md section 3, B-cpu): same structures, same arithmetic, weight-stationary threads */
#include "wn_cpu_best.h"
