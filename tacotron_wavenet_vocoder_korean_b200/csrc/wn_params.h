// wn_params.h -- structures shared by the host API (wn_api.cu) and the device code (wn_kernel.cu).
#pragma once
#include <stdint.h>
#include "../../include/wn_b200.h"

#define WN_NT 256                 // threads per CTA (8 warps), every role

// A matrix packed "thread-major" for one CTA: ncols dot products of length K.
// Thread tid handles column (pass*gpp + tid/t) and the contiguous k-chunk (tid % t) of length ch = K/t,
// which it evaluates as u consecutive sub-chains (canonical chunks (tid%t)*u .. (tid%t)*u + u-1 of length ch/u).
// Packed layout (V = 4):  w[((pass*(ch/4) + i4)*WN_NT + tid)*4 + j] = W[chunk*ch + 4*i4 + j][col]
//               (V = 1):  w[(pass*ch + i)*WN_NT + tid]              = W[chunk*ch + i][col]
// so that every warp-wide load is one fully coalesced, bank-conflict-free 512 B / 128 B access.
// The input vector is kept in shared memory with chunk c starting at c*xstride (xstride >= ch chosen
// so the t chunk starts fall in distinct banks).
struct WnMat {
    int32_t off;        // float offset of the packed matrix inside the CTA image
    int32_t K, ncols;
    int32_t t, ch, V;
    int32_t gpp;        // column groups per pass = WN_NT / t
    int32_t npass;
    int32_t xstride;
    int32_t xlen;       // t * xstride: floats of the padded input vector
    int32_t in_smem;    // 1: resident in shared memory, 0: read from the global image (overflow)
    int32_t u;          // independent fma sub-chains per thread; the canonical plan uses t*u chunks per column
};

// Shared-memory float offsets of the per-role scratch vectors (after the resident image prefix).
struct WnLayerSmem {
    int32_t xs_cur, xs_old, lcs, xraw, zs_dense, zs_skip, gvec, bfgN, pre, total_floats;
};
struct WnTailSmem {
    int32_t as1, c1s, total_floats;
};
struct WnSamplerSmem {
    int32_t c2s, cq, cqx, ids, qs, cdf /* doubles, 8B aligned */, red, misc, total_floats;
};

struct WnParams {
    // model dims
    int32_t N, L, R, D, S, O, Q, G, C, ifw, scalar_input, nr_mix;
    // topology
    int32_t M, Mt, Dm, Sm, St;           // Dm = D/M, Sm = S/M, St = S/Mt
    int32_t grid;
    // run
    int32_t T, n_forced, t_lc, lc_shift;
    float temperature;
    int32_t T_row[WN_MAX_BATCH];
    int32_t gc_id[WN_MAX_BATCH];
    int32_t dil[WN_MAX_LAYERS];
    // layer CTA image
    WnMat cur, old, lc, gc, dense, skip;
    int32_t off_bfg, off_bd, off_bs;
    int32_t layer_img_floats, layer_smem_floats;   // image size (stride) and resident prefix
    WnLayerSmem ls;
    // tail CTA image
    WnMat post1, post2;
    int32_t off_b1;
    int32_t tail_img_floats, tail_smem_floats;
    WnTailSmem ts;
    // sampler CTA image
    WnMat causal;
    int32_t off_b2;
    int32_t samp_img_floats, samp_smem_floats;
    WnSamplerSmem ss;
    // device pointers
    const float *layer_img, *tail_img, *samp_img;
    const float *gc_table;         // (card, G)
    const float *wc_onehot;        // (2, Q, R) causal kernel for one-hot input, read by row
    // Mailboxes live in a LOGICAL word space: these four are logical addresses (word index * 8, base 0),
    // never dereferenced.  A logical word w is stored at mb_tab[side][w >> 8] + (w & 255): 2 KB grains,
    // one copy homed in each die's L2 ("dual-homed").  Writers post to both copies, readers poll the copy
    // homed on their own die (sm_die[smid]).  Without calibration both tables point at the same grains.
    unsigned long long *mb_x;      // [N][L][M][R]    partial inputs of layer l
    unsigned long long *mb_z;      // [N][L][M][Dm]   gated activations, exchanged between the M siblings
    unsigned long long *mb_acc;    // [N][L][M][Sm]   running skip sum after layer l
    unsigned long long *mb_c2;     // [N][Mt][O]      partial conv2 outputs
    unsigned long long *const *mb_tab[2];
    const unsigned char *sm_die;   // [n_sm] die (0/1) of each SM id, or null
    int32_t mb_dual;               // 1: the two tables differ (post twice)
    int32_t pad2_;
    int32_t layer_base, layer_end; // cluster path: this launch runs layers [layer_base, layer_end)
    int32_t claim_slot;            // status[] index of this launch's die-aware cluster claim counters
    float *ring;                   // private dilation-queue rings
    const long long *ring_off;     // [L] float offset of layer l's ring block; block = [M][N][d][R]
    const float *forced;
    const float *lc_up;
    // folded create_upsample (cluster path): mel frames + the per-stage transposed-conv kernels; lc_up is then null
    const float *mel;              // (rows, t_mel, C) or null
    const float *upk;              // all stage kernels, stage s at up_off[s]: (F_s, 2)
    int32_t t_mel, n_up, hop;      // hop = prod(F_s)
    int32_t up_f[WN_MAX_UPSAMPLE], up_off[WN_MAX_UPSAMPLE];
    const void *uniforms;
    const float *noise;            // cluster path, scalar input: the draw's noise per (row, step): nr Gumbel values log(-log u), then the logistic
                                   // log u - log(1 - u), transformed from `uniforms` by wn_noise_prep_kernel before the launch
    float *out_samples;
    float *out_logits;
    int32_t *status;               // [0] abort flag, [1] cta that raised it, [2] code
    long long *prof;               // optional [grid][16] phase cycle counters (diagnostics), or null
};
