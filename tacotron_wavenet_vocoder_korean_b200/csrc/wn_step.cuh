// wn_step.cuh -- ONE network step with persistent device queues: the body of the reference's per-sample loop,
// sess.run(predict_proba_incremental) (generate.py:211, wavenet/model.py:215-245), for a caller that keeps that loop.
//
// State per row lives in global memory (wn_state): step counter, causal queue (model.py:122), the previous lc row
// (the lc queue of model.py:125 holds two rows, layers read the older one, model.py:79-80) and one ring per dilation
// queue (model.py:145) -- advanced by exactly one position per call, so a call costs O(1) in the number of steps taken.
// One CTA per row walks the layers with the weights read from the L2-resident TF-layout copy; every dot product goes
// through mv_plan_cta, the CTA-parallel twin of the CPU restatement's mv_plan (pinned arithmetic) with the SAME evaluation plan the persistent
// kernels implement, so a loop of wn_step calls reproduces wn_generate bit for bit.
// Included by wn_kernel.cu inside its anonymous namespace.
#pragma once

struct WnStepLayer {
    const float *wf, *wg, *bf, *bg, *gcf, *gcg, *lcf, *lcg, *wd, *bd, *ws, *bs;
};

struct WnStepParams {
    int32_t rows, L, R, D, S, O, Q, G, C, ifw, scalar, nr_mix;
    wn_plan plan;
    int32_t dil[WN_MAX_LAYERS];
    const WnStepLayer *layers;
    const float *wc, *w1, *b1, *w2, *b2, *gc_table;
    float *state;                    // [rows][state_stride]
    long long state_stride;          // floats per row
    int32_t off_cq, off_lc, off_ring0;   // float offsets inside a row's state: [0] step, [1] id_prev, [2] id_cur (as int bits)
    const long long *ring_off;       // [L] float offset of layer l's ring (d x R) from off_ring0
    const float *x_in;               // (rows) network input of this step (sample or mu-law id as float)
    const float *lc_row;             // (rows, C) the step's upsampled local-condition row, or null
    int32_t gc_id[WN_MAX_BATCH];
    const void *uniforms;            // (rows, nr_mix + 1) fp32 / (rows) fp64, or null: no draw
    float temperature;
    float *out_logits;               // (rows, out_dim) or null
    float *out_probs;                // one-hot models: (rows, Q) float32(softmax(float64(logits))) or null
    float *out_sample;               // (rows) drawn sample / id, or null
    int32_t smem_maxc;
};

// out[o] = sum_k W[k*stride + o] * x[k], o < ncols: t contiguous chunks, fma chain per chunk from +0, ascending xor-butterfly
__device__ void mv_plan_cta(const float *__restrict__ W, int stride, int ncols, const float *x, int K, int t, float *out, float *scratch)
{
    const int ch = K / t;
    for (int idx = threadIdx.x; idx < t * ncols; idx += blockDim.x) {
        const int c = idx / ncols, o = idx - c * ncols;
        const float *w = W + (size_t)(c * ch) * stride + o;
        const float *xc = x + c * ch;
        float a = 0.0f;
        for (int i = 0; i < ch; ++i) a = ffma(__ldg(w + (size_t)i * stride), xc[i], a);
        scratch[idx] = a;
    }
    __syncthreads();
    for (int off = 1; off < t; off <<= 1) {
        for (int idx = threadIdx.x; idx < t * ncols; idx += blockDim.x) {
            const int c = idx / ncols, o = idx - c * ncols;
            if ((c & off) == 0) {
                const float v = fadd(scratch[idx], scratch[(c ^ off) * ncols + o]);
                scratch[idx] = v;
                scratch[(c ^ off) * ncols + o] = v;
            }
        }
        __syncthreads();
    }
    for (int o = threadIdx.x; o < ncols; o += blockDim.x) out[o] = scratch[o];
    __syncthreads();
}

__global__ void __launch_bounds__(WN_NT, 1) wn_step_kernel(const __grid_constant__ WnStepParams p)
{
    const int b = blockIdx.x, tid = threadIdx.x;
    const int L = p.L, R = p.R, D = p.D, S = p.S, O = p.O, Q = p.Q, G = p.G, C = p.C, ifw = p.ifw;
    const int maxc = p.smem_maxc;
    float *sm = g_smem;
    float *x = sm, *xn = x + R, *z = xn + R, *f = z + D, *g = f + D, *acc = g + D, *c1 = acc + S, *c2 = c1 + S;
    float *tmp = c2 + (((O > WN_NT ? O : WN_NT) + 3) & ~3), *gvec = tmp + maxc, *lcp = gvec + ((G + 3) & ~3), *cq = lcp + ((C + 3) & ~3);
    float *misc = cq + ((ifw + 3) & ~3);
    double *red = reinterpret_cast<double *>(misc + 32);
    double *cdf = red + 16;
    float *scratch = reinterpret_cast<float *>(cdf + 2);
    float *st = p.state + (size_t)b * p.state_stride;
    int *sti = reinterpret_cast<int *>(st);
    const int t = sti[0];
    const float x_in = __ldg(p.x_in + b);

    // ---- queues in, causal layer (model.py:122,131) ------------------------------------------------------------
    if (p.scalar) {
        for (int i = tid; i < ifw; i += blockDim.x) cq[i] = (i < ifw - 1) ? st[p.off_cq + i + 1] : x_in;
    }
    for (int i = tid; i < C; i += blockDim.x) lcp[i] = st[p.off_lc + i];
    if (G) for (int i = tid; i < G; i += blockDim.x) gvec[i] = __ldg(p.gc_table + (size_t)p.gc_id[b] * G + i);
    __syncthreads();
    if (p.scalar) {
        for (int i = tid; i < ifw; i += blockDim.x) st[p.off_cq + i] = cq[i];
        mv_plan_cta(p.wc, R, R, cq, ifw, p.plan.t_causal, x, scratch);
    } else {
        const int prev = sti[2], cur = (int)x_in;                   // id_cur of the previous step becomes id_prev
        for (int r = tid; r < R; r += blockDim.x) {
            const float a = (prev >= 0) ? __ldg(p.wc + ((size_t)0 * Q + prev) * R + r) : 0.0f;
            const float bb = (cur >= 0 && cur < Q) ? __ldg(p.wc + ((size_t)1 * Q + cur) * R + r) : 0.0f;
            x[r] = fadd(a, bb);
        }
        __syncthreads();
        if (tid == 0) { sti[1] = prev; sti[2] = cur; }
    }
    // ---- dilated stack (model.py:141-149, 66-101) --------------------------------------------------------------
    for (int l = 0; l < L; ++l) {
        const WnStepLayer ly = p.layers[l];
        const int d = p.dil[l];
        float *slot = st + p.off_ring0 + p.ring_off[l] + (size_t)(t % d) * R;      // holds x_l(t - d)
        for (int o = tid; o < D; o += blockDim.x) { f[o] = __ldg(ly.bf + o); g[o] = __ldg(ly.bg + o); }
        __syncthreads();
        if (G) {
            mv_plan_cta(ly.gcf, D, D, gvec, G, p.plan.t_gc, tmp, scratch);
            for (int o = tid; o < D; o += blockDim.x) f[o] = fadd(f[o], tmp[o]);
            __syncthreads();
            mv_plan_cta(ly.gcg, D, D, gvec, G, p.plan.t_gc, tmp, scratch);
            for (int o = tid; o < D; o += blockDim.x) g[o] = fadd(g[o], tmp[o]);
            __syncthreads();
        }
        for (int r = tid; r < R; r += blockDim.x) xn[r] = slot[r];                  // dilated tap
        __syncthreads();
        mv_plan_cta(ly.wf, D, D, xn, R, p.plan.t_old, tmp, scratch);
        for (int o = tid; o < D; o += blockDim.x) f[o] = fadd(f[o], tmp[o]);
        __syncthreads();
        mv_plan_cta(ly.wg, D, D, xn, R, p.plan.t_old, tmp, scratch);
        for (int o = tid; o < D; o += blockDim.x) g[o] = fadd(g[o], tmp[o]);
        __syncthreads();
        if (C) {
            mv_plan_cta(ly.lcf, D, D, lcp, C, p.plan.t_lc, tmp, scratch);
            for (int o = tid; o < D; o += blockDim.x) f[o] = fadd(f[o], tmp[o]);
            __syncthreads();
            mv_plan_cta(ly.lcg, D, D, lcp, C, p.plan.t_lc, tmp, scratch);
            for (int o = tid; o < D; o += blockDim.x) g[o] = fadd(g[o], tmp[o]);
            __syncthreads();
        }
        mv_plan_cta(ly.wf + (size_t)R * D, D, D, x, R, p.plan.t_cur, tmp, scratch);
        for (int o = tid; o < D; o += blockDim.x) f[o] = fadd(f[o], tmp[o]);
        __syncthreads();
        mv_plan_cta(ly.wg + (size_t)R * D, D, D, x, R, p.plan.t_cur, tmp, scratch);
        for (int o = tid; o < D; o += blockDim.x) z[o] = fmul(wn::tanh32(f[o]), wn::sigmoid32(fadd(g[o], tmp[o])));   // model.py:86
        for (int r = tid; r < R; r += blockDim.x) slot[r] = x[r];                   // queue push (model.py:145)
        __syncthreads();
        mv_plan_cta(ly.ws, S, S, z, D, p.plan.t_skip, tmp, scratch);
        for (int s = tid; s < S; s += blockDim.x) {
            const float v = fadd(__ldg(ly.bs + s), tmp[s]);
            acc[s] = (l == 0) ? v : fadd(acc[s], v);                                  // sum(outputs), model.py:157
        }
        for (int r = tid; r < R; r += blockDim.x) xn[r] = fadd(x[r], __ldg(ly.bd + r));
        __syncthreads();
        const int dm = D / p.plan.M;
        for (int mm = 0; mm < p.plan.M; ++mm) {
            mv_plan_cta(ly.wd + (size_t)mm * dm * R, R, R, z + mm * dm, dm, p.plan.t_dense, tmp, scratch);
            for (int r = tid; r < R; r += blockDim.x) xn[r] = fadd(xn[r], tmp[r]);
            __syncthreads();
        }
        for (int r = tid; r < R; r += blockDim.x) x[r] = xn[r];
        __syncthreads();
    }
    // ---- postprocessing (model.py:150-165) ----------------------------------------------------------------------
    for (int s = tid; s < S; s += blockDim.x) acc[s] = relu32(acc[s]);
    __syncthreads();
    mv_plan_cta(p.w1, S, S, acc, S, p.plan.t_post1, tmp, scratch);
    for (int s = tid; s < S; s += blockDim.x) c1[s] = relu32(fadd(__ldg(p.b1 + s), tmp[s]));
    for (int o = tid; o < O; o += blockDim.x) c2[o] = __ldg(p.b2 + o);
    __syncthreads();
    const int sm_ = S / p.plan.Mt;
    for (int mm = 0; mm < p.plan.Mt; ++mm) {
        mv_plan_cta(p.w2 + (size_t)mm * sm_ * O, O, O, c1 + mm * sm_, sm_, p.plan.t_post2, tmp, scratch);
        for (int o = tid; o < O; o += blockDim.x) c2[o] = fadd(c2[o], tmp[o]);
        __syncthreads();
    }
    if (p.out_logits) for (int o = tid; o < O; o += blockDim.x) p.out_logits[(size_t)b * O + o] = c2[o];
    // ---- head (model.py:238-243) + draw (mixture.py:84-114 / generate.py:219-231) ----------------------------------
    if (p.scalar) {
        if (p.uniforms != nullptr && tid < 32) {
            const int nr = p.nr_mix, lane = tid;
            const float *u = (const float *)p.uniforms + (size_t)b * (nr + 1);
            float gq = (lane < nr) ? fsub(c2[lane], wn::log32(-wn::log32(__ldg(u + lane)))) : __int_as_float(0xff800000);
            int k = lane;
            for (int off = 1; off < 32; off <<= 1) {
                const float og = __shfl_xor_sync(FULL, gq, off);
                const int ok = __shfl_xor_sync(FULL, k, off);
                if (og > gq || (og == gq && ok < k)) { gq = og; k = ok; }
            }
            if (lane == 0) {
                const float u2 = __ldg(u + nr);
                const float logistic = fsub(wn::log32(u2), wn::log32(fsub(1.0f, u2)));
                float ls = c2[2 * nr + k];
                const float lsmin = -32.23619130191664f;
                if (!(ls > lsmin)) ls = lsmin;
                float xs = fadd(c2[nr + k], fmul(wn::exp32(ls), logistic));
                xs = fminf(fmaxf(xs, -1.0f), 1.0f);
                if (p.out_sample) p.out_sample[b] = xs;
            }
        }
    } else {
        const double u = p.uniforms ? __ldg((const double *)p.uniforms + b) : 0.5;
        const float temp = p.uniforms ? p.temperature : 1.0f;
        const float id = mulaw_draw_cta(c2, Q, temp, u, misc, red, cdf, p.out_probs ? p.out_probs + (size_t)b * Q : nullptr);
        if (p.out_sample && p.uniforms && tid == 0) p.out_sample[b] = id;
    }
    // ---- lc queue push (model.py:125) and step counter ----------------------------------------------------------------
    for (int i = tid; i < C; i += blockDim.x) st[p.off_lc + i] = p.lc_row ? __ldg(p.lc_row + (size_t)b * C + i) : 0.0f;
    if (tid == 0) sti[0] = t + 1;
}
