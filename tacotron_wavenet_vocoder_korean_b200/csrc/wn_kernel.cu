// wn_kernel.cu -- the B200 (sm_100a) persistent WaveNet sample-loop kernel.
//
// One cooperative launch generates every audio sample of every utterance in the batch
// (the hot loop of generate.py:202-233 around predict_proba_incremental, wavenet/model.py:215-245).
//
// Mapping (DESIGN.md "Kernel"):
//   * layer CTAs  (l, m), l < L, m < M : own 1/M of layer l's weights, resident in shared memory
//       - fg columns of the dilated filter/gate convs for D/M gated channels (model.py:68-83)
//       - the matching K-slice of the dense 1x1 (model.py:89) -> partial residual outputs
//       - S/M columns of the skip 1x1 (model.py:96) and the running skip sum (model.py:157)
//   * tail CTAs   mt < Mt : S/Mt columns of postprocess conv1 and the matching K-slice of conv2
//       (model.py:158-165)
//   * one sampler CTA: conv2 reduction, MoL / mu-law draw (mixture.py:84-114, generate.py:219-231),
//       causal queue + causal conv (model.py:41-46,122,131), feeds layer 0.
//   Activations hop CTA -> CTA through L2 "LL" mailboxes: 8-byte words {fp32 value, step tag}
//   written with st.relaxed.gpu and polled with ld.relaxed.gpu -- no fences, no grid barrier.
//   Utterances are software-pipelined through the layer chain (row b is in layer l while row b+1 is
//   in layer l-1).
//
// Arithmetic follows DESIGN.md "Pinned arithmetic": every dot product is evaluated in the order the
// plan (wn_get_plan) describes, every transcendental through wn_math.cuh, so results are a pure
// function of the inputs and can be compared bit-for-bit with the CPU oracle.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "wn_params.h"
#include "wn_math.cuh"

namespace {

using wn::fadd;
using wn::fsub;
using wn::fmul;
using wn::fdiv;
using wn::ffma;

typedef unsigned long long u64;

constexpr unsigned FULL = 0xffffffffu;
constexpr long long WATCHDOG_CYCLES = 3000000000LL;   // ~1.5 s of a single stalled wait

// ---------------------------------------------------------------------------------------------
// LL mailboxes
__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ll_post(u64 *p, float val, unsigned seq)
{
    u64 v = ((u64)seq << 32) | (u64)__float_as_uint(val);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_i32(const int *p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct Abort {
    int32_t *status;
    int flag;
};

__device__ __noinline__ bool watchdog_check(Abort &ab, long long &t0)
{
    if (t0 == 0) t0 = clock64();
    if (ld_volatile_i32(ab.status) != 0) { ab.flag = 1; return true; }
    if (clock64() - t0 > WATCHDOG_CYCLES) {
        if (atomicCAS(ab.status, 0, 1) == 0) { ab.status[1] = (int)blockIdx.x; ab.status[2] = (int)threadIdx.x; }
        ab.flag = 1;
        return true;
    }
    return false;
}

__device__ __forceinline__ float ll_wait(const u64 *p, unsigned seq, Abort &ab)
{
    u64 v = ld_relaxed_u64(p);
    unsigned spins = 0;
    long long t0 = 0;
    while ((unsigned)(v >> 32) != seq) {
        if (((++spins) & 0x3ffu) == 0 && watchdog_check(ab, t0)) break;
        v = ld_relaxed_u64(p);
    }
    return __uint_as_float((unsigned)v);
}

// wait for n (<=4) words p[i*stride]; loads are issued together so the latencies overlap
__device__ __forceinline__ void ll_wait_n(const u64 *p, size_t stride, int n, unsigned seq, Abort &ab, float *out)
{
    u64 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (i < n) ? ld_relaxed_u64(p + i * stride) : ((u64)seq << 32);
    unsigned spins = 0;
    long long t0 = 0;
    while (true) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 4; ++i) ok = ok && ((unsigned)(v[i] >> 32) == seq);
        if (ok) break;
        if (((++spins) & 0x3ffu) == 0 && watchdog_check(ab, t0)) break;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < n && (unsigned)(v[i] >> 32) != seq) v[i] = ld_relaxed_u64(p + i * stride);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = __uint_as_float((unsigned)v[i]);
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copy of the resident weight image (global -> shared), completion on an mbarrier.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ void load_image_tma(float *smem_dst, const float *gsrc, int n_floats, uint64_t *bar)
{
    const int tid = threadIdx.x;
    const uint32_t bytes = (uint32_t)n_floats * 4u;
    if (bytes == 0) return;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        const uint32_t CH = 32768u;
        for (uint32_t o = 0; o < bytes; o += CH) {
            uint32_t sz = (bytes - o < CH) ? (bytes - o) : CH;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32((const char *)smem_dst + o)),
                "l"((const char *)gsrc + o), "r"(sz), "r"(smem_u32(bar))
                : "memory");
        }
    }
    // every thread waits for phase 0
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar))
            : "memory");
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Thread-major matvec.  All WN_NT threads must call it (warp shuffles inside).
__device__ __forceinline__ int xpad(const WnMat &m, int k) { return (k / m.ch) * m.xstride + (k % m.ch); }

template <class F>
__device__ __forceinline__ void matvec(const WnMat &m, const float *__restrict__ w, const float *__restrict__ xs, F &&epi)
{
    const int tid = threadIdx.x;
    const int chunk = tid % m.t;
    const int grp = tid / m.t;
    const float *xc = xs + chunk * m.xstride;
    for (int pass = 0; pass < m.npass; ++pass) {
        float acc = 0.0f;
        if (m.V == 4) {
            const float4 *w4 = reinterpret_cast<const float4 *>(w) + (size_t)pass * (m.ch >> 2) * WN_NT + tid;
            const float4 *x4 = reinterpret_cast<const float4 *>(xc);
            const int n4 = m.ch >> 2;
#pragma unroll 4
            for (int i = 0; i < n4; ++i) {
                float4 wv = w4[(size_t)i * WN_NT];
                float4 xv = x4[i];
                acc = ffma(wv.x, xv.x, acc);
                acc = ffma(wv.y, xv.y, acc);
                acc = ffma(wv.z, xv.z, acc);
                acc = ffma(wv.w, xv.w, acc);
            }
        } else {
            const float *w1 = w + (size_t)pass * m.ch * WN_NT + tid;
#pragma unroll 4
            for (int i = 0; i < m.ch; ++i) acc = ffma(w1[(size_t)i * WN_NT], xc[i], acc);
        }
        for (int off = 1; off < m.t; off <<= 1) acc = fadd(acc, __shfl_xor_sync(FULL, acc, off));
        const int col = pass * m.gpp + grp;
        if (chunk == 0 && col < m.ncols) epi(col, acc);
    }
}

// tanh (is_gate = false) or sigmoid (is_gate = true) without divergence; the operation sequences are
// exactly those of wn::tanh32 / wn::sigmoid32.
__device__ __forceinline__ float act_fg(float x, bool is_gate)
{
    float ax = fabsf(x);
    float arg = is_gate ? -x : fadd(ax, ax);
    float e = wn::exp32(arg);
    float q = fdiv(is_gate ? 1.0f : 2.0f, fadd(e, 1.0f));
    float th = (ax > 44.0f) ? 1.0f : fsub(1.0f, q);
    return is_gate ? q : copysignf(th, x);
}

__device__ __forceinline__ float relu32(float v) { return v > 0.0f ? v : 0.0f; }

// =============================================================================================
// Layer CTA
__device__ void layer_role(const WnParams &p, float *smem, int l, int m)
{
    const int tid = threadIdx.x;
    const int N = p.N, L = p.L, R = p.R, M = p.M, Dm = p.Dm, Sm = p.Sm;
    const int cta = l * M + m;
    const float *gimg = p.layer_img + (size_t)cta * p.layer_img_floats;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.layer_smem_floats, &bar);

    auto wptr = [&](const WnMat &mt) -> const float * { return mt.in_smem ? (smem + mt.off) : (gimg + mt.off); };
    const float *w_cur = wptr(p.cur), *w_old = wptr(p.old), *w_lc = wptr(p.lc), *w_gc = wptr(p.gc);
    const float *w_dense = wptr(p.dense), *w_skip = wptr(p.skip);
    // small vectors are always inside the resident prefix
    const float *bfg = smem + p.off_bfg, *bd = smem + p.off_bd, *bs = smem + p.off_bs;

    float *sc = smem + p.layer_smem_floats;
    float *xs_cur = sc + p.ls.xs_cur, *xs_old = sc + p.ls.xs_old, *lcs = sc + p.ls.lcs, *xraw = sc + p.ls.xraw;
    float *zs_dense = sc + p.ls.zs_dense, *zs_skip = sc + p.ls.zs_skip, *gvec = sc + p.ls.gvec;
    float *bfgN = sc + p.ls.bfgN, *pre = sc + p.ls.pre;

    const int d = p.dil[l];
    const int nin = (l == 0) ? 1 : M;
    const int ncol2 = 2 * Dm;
    float *ring_cta = p.ring + p.ring_off[l] + (size_t)m * N * d * R;
    Abort ab{p.status, 0};

    // zero the padded vectors once (pad lanes are never read, but keep them defined)
    for (int i = tid; i < p.cur.xlen; i += WN_NT) xs_cur[i] = 0.0f;
    for (int i = tid; i < p.old.xlen; i += WN_NT) xs_old[i] = 0.0f;
    for (int i = tid; i < p.lc.xlen; i += WN_NT) lcs[i] = 0.0f;
    for (int i = tid; i < p.gc.xlen; i += WN_NT) gvec[i] = 0.0f;
    for (int i = tid; i < p.dense.xlen; i += WN_NT) zs_dense[i] = 0.0f;
    for (int i = tid; i < p.skip.xlen; i += WN_NT) zs_skip[i] = 0.0f;
    __syncthreads();

    // pre-activation for step tn of row b: bias(+gc) + W_old . x_l(tn-d) + W_lc . lc(tn-1)   (off the chain)
    auto compute_pre = [&](int b, int tn) {
        if (tid < R) {
            float v;
            if (d == 1) v = (tn == 0) ? 0.0f : xraw[tid];
            else v = __ldcg(ring_cta + ((size_t)b * d + (tn % d)) * R + tid);
            xs_old[xpad(p.old, tid)] = v;
        }
        if (p.C && tid < p.C) {
            long idx = (long)tn - 1 - p.lc_shift;
            float v = 0.0f;
            if (p.lc_up != nullptr && idx >= 0 && idx < p.t_lc) v = __ldg(p.lc_up + ((size_t)b * p.t_lc + idx) * p.C + tid);
            lcs[xpad(p.lc, tid)] = v;
        }
        __syncthreads();
        matvec(p.old, w_old, xs_old, [&](int col, float dot) { pre[b * ncol2 + col] = fadd(bfgN[b * ncol2 + col], dot); });
        if (p.C) {
            __syncthreads();
            matvec(p.lc, w_lc, lcs, [&](int col, float dot) { pre[b * ncol2 + col] = fadd(pre[b * ncol2 + col], dot); });
        }
        __syncthreads();
    };

    // ---- prologue: fold the speaker embedding into the biases (model.py:71-73,181-212), pre for t = 0
    for (int b = 0; b < N; ++b) {
        if (p.G) {
            if (tid < p.G) gvec[xpad(p.gc, tid)] = __ldg(p.gc_table + (size_t)p.gc_id[b] * p.G + tid);
            __syncthreads();
            matvec(p.gc, w_gc, gvec, [&](int col, float dot) { bfgN[b * ncol2 + col] = fadd(bfg[col], dot); });
        } else {
            if (tid < ncol2) bfgN[b * ncol2 + tid] = bfg[tid];
        }
        __syncthreads();
        compute_pre(b, 0);
    }

    // ---- main loop
    const int grp = tid / p.cur.t, chunk = tid % p.cur.t;
    for (int t = 0; t < p.T; ++t) {
        const unsigned seq = (unsigned)t + 1u;
        for (int b = 0; b < N; ++b) {
            if (t >= p.T_row[b]) continue;
            // 1. wait for the layer input: sum of the partial residual outputs of layer l-1
            if (tid < R) {
                float q[4];
                ll_wait_n(p.mb_x + (((size_t)b * L + l) * M) * R + tid, (size_t)R, nin, seq, ab, q);
                float v = q[0];
                for (int i = 1; i < nin; ++i) v = fadd(v, q[i]);
                xs_cur[xpad(p.cur, tid)] = v;
                xraw[tid] = v;
            }
            if (__syncthreads_or(ab.flag)) return;

            // 2. filter/gate for the current tap + gated activation (model.py:68-69,86)
            {
                float acc = 0.0f;
                const float *xc = xs_cur + chunk * p.cur.xstride;
                if (p.cur.V == 4) {
                    const float4 *w4 = reinterpret_cast<const float4 *>(w_cur) + tid;
                    const float4 *x4 = reinterpret_cast<const float4 *>(xc);
                    const int n4 = p.cur.ch >> 2;
#pragma unroll 8
                    for (int i = 0; i < n4; ++i) {
                        float4 wv = w4[(size_t)i * WN_NT];
                        float4 xv = x4[i];
                        acc = ffma(wv.x, xv.x, acc);
                        acc = ffma(wv.y, xv.y, acc);
                        acc = ffma(wv.z, xv.z, acc);
                        acc = ffma(wv.w, xv.w, acc);
                    }
                } else {
                    const float *w1 = w_cur + tid;
                    for (int i = 0; i < p.cur.ch; ++i) acc = ffma(w1[(size_t)i * WN_NT], xc[i], acc);
                }
                for (int off = 1; off < p.cur.t; off <<= 1) acc = fadd(acc, __shfl_xor_sync(FULL, acc, off));
                const bool valid = grp < ncol2;
                float pv = valid ? pre[b * ncol2 + grp] : 0.0f;
                float fg = fadd(pv, acc);
                const bool is_gate = (grp & 1) != 0;
                float a = act_fg(fg, is_gate);
                float other = __shfl_xor_sync(FULL, a, p.cur.t);   // partner column (filter <-> gate)
                if (valid && chunk == 0 && !is_gate) {
                    float z = fmul(a, other);
                    int j = grp >> 1;
                    zs_dense[xpad(p.dense, j)] = z;
                    zs_skip[xpad(p.skip, m * Dm + j)] = z;
                    if (M > 1) ll_post(p.mb_z + (((size_t)b * L + l) * M + m) * Dm + j, z, seq);
                }
            }
            __syncthreads();

            // 3. partial dense 1x1 + residual (model.py:89,98-101) -> mailbox of layer l+1
            if (l + 1 < L) {
                u64 *dst = p.mb_x + (((size_t)b * L + (l + 1)) * M + m) * R;
                matvec(p.dense, w_dense, zs_dense, [&](int r, float dot) {
                    float v = (m == 0) ? fadd(fadd(xraw[r], bd[r]), dot) : dot;
                    ll_post(dst + r, v, seq);
                });
            }
            // ---- everything below is off the sample-to-sample critical chain ----
            // 4. push x_l(t) into the private dilation-queue ring (model.py:145)
            if (d >= 2 && tid < R) __stcg(ring_cta + ((size_t)b * d + (t % d)) * R + tid, xraw[tid]);
            // 5. gather the sibling CTAs' gated activations
            if (M > 1 && tid < p.D) {
                int mm = tid / Dm;
                if (mm != m) {
                    float z = ll_wait(p.mb_z + (((size_t)b * L + l) * M + mm) * Dm + (tid % Dm), seq, ab);
                    zs_skip[xpad(p.skip, tid)] = z;
                }
            }
            if (__syncthreads_or(ab.flag)) return;
            // 6. skip 1x1 (model.py:94-96) + running sum over layers (model.py:157)
            {
                const u64 *src = (l > 0) ? p.mb_acc + (((size_t)b * L + (l - 1)) * M + m) * Sm : nullptr;
                u64 *dst = p.mb_acc + (((size_t)b * L + l) * M + m) * Sm;
                matvec(p.skip, w_skip, zs_skip, [&](int c, float dot) {
                    float v = fadd(bs[c], dot);
                    if (l > 0) v = fadd(ll_wait(src + c, seq, ab), v);
                    ll_post(dst + c, v, seq);
                });
            }
            // 7. pre-activations of the next step
            if (t + 1 < p.T_row[b]) compute_pre(b, t + 1);
            else __syncthreads();
            if (__syncthreads_or(ab.flag)) return;
        }
    }
}

// =============================================================================================
// Tail CTA: relu -> conv1 (S->S) -> relu -> partial conv2 (model.py:158-165)
__device__ void tail_role(const WnParams &p, float *smem, int mt)
{
    const int tid = threadIdx.x;
    const int N = p.N, L = p.L, M = p.M, Sm = p.Sm, S = p.S;
    const float *gimg = p.tail_img + (size_t)mt * p.tail_img_floats;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.tail_smem_floats, &bar);
    const float *w1 = p.post1.in_smem ? smem + p.post1.off : gimg + p.post1.off;
    const float *w2 = p.post2.in_smem ? smem + p.post2.off : gimg + p.post2.off;
    const float *b1 = smem + p.off_b1;
    float *sc = smem + p.tail_smem_floats;
    float *as1 = sc + p.ts.as1, *c1s = sc + p.ts.c1s;
    for (int i = tid; i < p.post1.xlen; i += WN_NT) as1[i] = 0.0f;
    for (int i = tid; i < p.post2.xlen; i += WN_NT) c1s[i] = 0.0f;
    __syncthreads();
    Abort ab{p.status, 0};
    for (int t = 0; t < p.T; ++t) {
        const unsigned seq = (unsigned)t + 1u;
        for (int b = 0; b < N; ++b) {
            if (t >= p.T_row[b]) continue;
            const u64 *src = p.mb_acc + (((size_t)b * L + (L - 1)) * M) * Sm;   // [M][Sm] == S contiguous words
            for (int c = tid; c < S; c += WN_NT) as1[xpad(p.post1, c)] = relu32(ll_wait(src + c, seq, ab));
            if (__syncthreads_or(ab.flag)) return;
            matvec(p.post1, w1, as1, [&](int c, float dot) { c1s[xpad(p.post2, c)] = relu32(fadd(b1[c], dot)); });
            __syncthreads();
            u64 *dst = p.mb_c2 + ((size_t)b * p.Mt + mt) * p.O;
            matvec(p.post2, w2, c1s, [&](int o, float dot) { ll_post(dst + o, dot, seq); });
            __syncthreads();
        }
    }
}

// =============================================================================================
// Sampler CTA
__device__ void sampler_role(const WnParams &p, float *smem)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, R = p.R, O = p.O, Q = p.Q, nr = p.nr_mix, ifw = p.ifw;
    const float *gimg = p.samp_img;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.samp_smem_floats, &bar);
    const float *w_c = p.causal.in_smem ? smem + p.causal.off : gimg + p.causal.off;
    const float *b2 = smem + p.off_b2;
    float *sc = smem + p.samp_smem_floats;
    float *c2s = sc + p.ss.c2s, *cq = sc + p.ss.cq, *cqx = sc + p.ss.cqx, *qs = sc + p.ss.qs;
    int *ids = reinterpret_cast<int *>(sc + p.ss.ids);
    double *cdf = reinterpret_cast<double *>(sc + p.ss.cdf);
    double *red = reinterpret_cast<double *>(sc + p.ss.red);     // 16 doubles
    float *misc = sc + p.ss.misc;                                 // [0] sample
    Abort ab{p.status, 0};

    for (int i = tid; i < N * ifw; i += WN_NT) cq[i] = 0.0f;
    for (int i = tid; i < p.causal.xlen; i += WN_NT) cqx[i] = 0.0f;
    for (int i = tid; i < 2 * N; i += WN_NT) ids[i] = -1;
    __syncthreads();

    // push x_in into row b's causal queue, run the causal conv, post to layer 0 with tag seq
    auto feed = [&](int b, float x_in, unsigned seq) {
        u64 *dst = p.mb_x + (((size_t)b * p.L + 0) * p.M + 0) * R;
        if (p.scalar_input) {
            float v = 0.0f;
            if (tid < ifw) v = (tid < ifw - 1) ? cq[b * ifw + tid + 1] : x_in;
            __syncthreads();
            if (tid < ifw) { cq[b * ifw + tid] = v; cqx[xpad(p.causal, tid)] = v; }
            __syncthreads();
            matvec(p.causal, w_c, cqx, [&](int r, float dot) { ll_post(dst + r, dot, seq); });
        } else {
            int prev = ids[2 * b + 1];
            int cur = (int)x_in;
            __syncthreads();
            if (tid == 0) { ids[2 * b] = prev; ids[2 * b + 1] = cur; }
            if (tid < R) {
                float a = (prev >= 0) ? __ldg(p.wc_onehot + ((size_t)0 * Q + prev) * R + tid) : 0.0f;
                float bb = (cur >= 0 && cur < Q) ? __ldg(p.wc_onehot + ((size_t)1 * Q + cur) * R + tid) : 0.0f;
                ll_post(dst + tid, fadd(a, bb), seq);
            }
        }
        __syncthreads();
    };

    for (int b = 0; b < N; ++b)
        if (p.T_row[b] > 0) feed(b, __ldg(p.forced + (size_t)b * p.n_forced), 1u);

    for (int t = 0; t < p.T; ++t) {
        const unsigned seq = (unsigned)t + 1u;
        for (int b = 0; b < N; ++b) {
            if (t >= p.T_row[b]) continue;
            // prefetch this step's uniforms and the next forced input while the network runs
            float gum = 0.0f, logistic = 0.0f, next_forced = 0.0f;
            double u64v = 0.0;
            if (p.scalar_input) {
                const float *u = (const float *)p.uniforms + ((size_t)b * p.T + t) * (nr + 1);
                if (warp == 0 && lane < nr) gum = wn::log32(-wn::log32(__ldg(u + lane)));
                if (tid == 0) { float u2 = __ldg(u + nr); logistic = fsub(wn::log32(u2), wn::log32(fsub(1.0f, u2))); }
            } else {
                u64v = __ldg((const double *)p.uniforms + (size_t)b * p.T + t);
            }
            const bool has_next = (t + 1 < p.T_row[b]);
            if (has_next && t + 1 < p.n_forced) next_forced = __ldg(p.forced + (size_t)b * p.n_forced + t + 1);

            // conv2 output: bias + the Mt partial sums in order
            if (tid < O) {
                float v = b2[tid];
                const u64 *src = p.mb_c2 + ((size_t)b * p.Mt) * O + tid;
                for (int m0 = 0; m0 < p.Mt; m0 += 4) {
                    float q[4];
                    int n = (p.Mt - m0 < 4) ? (p.Mt - m0) : 4;
                    ll_wait_n(src + (size_t)m0 * O, (size_t)O, n, seq, ab, q);
                    for (int i = 0; i < n; ++i) v = fadd(v, q[i]);
                }
                c2s[tid] = v;
                if (p.out_logits) p.out_logits[((size_t)b * p.T + t) * O + tid] = v;
            }
            if (__syncthreads_or(ab.flag)) return;

            float sample;
            if (p.scalar_input) {
                // discretized mixture of logistics draw, wavenet/mixture.py:84-114
                if (warp == 0) {
                    float g = (lane < nr) ? fsub(c2s[lane], gum) : __int_as_float(0xff800000);
                    int k = lane;
                    for (int off = 1; off < 32; off <<= 1) {
                        float og = __shfl_xor_sync(FULL, g, off);
                        int ok = __shfl_xor_sync(FULL, k, off);
                        if (og > g || (og == g && ok < k)) { g = og; k = ok; }
                    }
                    if (lane == 0) {
                        float mean = c2s[nr + k];
                        float ls = c2s[2 * nr + k];
                        const float lsmin = -32.23619130191664f;
                        if (!(ls > lsmin)) ls = lsmin;
                        float x = fadd(mean, fmul(wn::exp32(ls), logistic));
                        x = fmaxf(x, -1.0f);
                        x = fminf(x, 1.0f);
                        misc[0] = x;
                    }
                }
                __syncthreads();
                sample = misc[0];
            } else {
                // float64 softmax (model.py:243) -> fp32; temperature + categorical draw (generate.py:219-231)
                const int nw = Q >> 5;
                const bool act = tid < Q;
                float c = act ? c2s[tid] : __int_as_float(0xff800000);
                float mx = c;
                for (int off = 1; off < 32; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, off));
                if (lane == 0) misc[8 + warp] = mx;
                __syncthreads();
                mx = misc[8];
                for (int w = 1; w < nw; ++w) mx = fmaxf(mx, misc[8 + w]);
                double e = act ? wn::exp64((double)c - (double)mx) : 0.0;
                double sum = e;
                for (int off = 1; off < 32; off <<= 1) sum = __dadd_rn(sum, __shfl_xor_sync(FULL, sum, off));
                if (lane == 0 && act) red[warp] = sum;
                __syncthreads();
                double arr[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) arr[w] = (w < nw) ? red[w] : 0.0;
                for (int off = 1; off < nw; off <<= 1) {
                    double nx[8];
#pragma unroll
                    for (int w = 0; w < 8; ++w) nx[w] = (w < nw) ? __dadd_rn(arr[w], arr[(w ^ off) & 7]) : 0.0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) arr[w] = nx[w];
                }
                const double den = arr[0];
                float pr = act ? (float)__ddiv_rn(e, den) : 0.0f;
                float s = fdiv(wn::log32(pr), p.temperature);
                float a = s;
                for (int off = 1; off < 32; off <<= 1) {
                    float o = __shfl_xor_sync(FULL, a, off);
                    a = (lane & off) ? wn::logaddexp32(o, a) : wn::logaddexp32(a, o);
                }
                __syncthreads();                     // red/misc reuse
                if (lane == 0 && act) misc[16 + warp] = a;
                __syncthreads();
                float fa[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) fa[w] = (w < nw) ? misc[16 + w] : 0.0f;
                for (int off = 1; off < nw; off <<= 1) {
                    float nx[8];
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        int o = (w ^ off) & 7;
                        nx[w] = (w < nw) ? ((w < o) ? wn::logaddexp32(fa[w], fa[o]) : wn::logaddexp32(fa[o], fa[w])) : 0.0f;
                    }
#pragma unroll
                    for (int w = 0; w < 8; ++w) fa[w] = nx[w];
                }
                const float lse = fa[0];
                if (act) qs[tid] = wn::exp32(fsub(s, lse));
                __syncthreads();
                if (tid == 0) {
                    double accd = 0.0;
                    for (int j = 0; j < Q; ++j) { accd = __dadd_rn(accd, (double)qs[j]); cdf[j] = accd; }   // np.cumsum
                    red[8] = accd;
                }
                __syncthreads();
                const double total = red[8];
                int pred = act && (__ddiv_rn(cdf[tid < Q ? tid : 0], total) <= u64v);
                int cnt = __syncthreads_count(pred);
                if (cnt > Q - 1) cnt = Q - 1;
                sample = (float)cnt;
            }
            if (tid == 0) p.out_samples[(size_t)b * p.T + t] = sample;
            if (has_next) feed(b, (t + 1 < p.n_forced) ? next_forced : sample, seq + 1u);
            else __syncthreads();
        }
    }
}

}  // namespace

// =============================================================================================
extern "C" __global__ void __launch_bounds__(WN_NT, 1) wn_persistent_kernel(const __grid_constant__ WnParams p)
{
    extern __shared__ __align__(128) float smem[];
    const int cta = blockIdx.x;
    const int n_layer = p.L * p.M;
    if (cta < n_layer) layer_role(p, smem, cta / p.M, cta % p.M);
    else if (cta < n_layer + p.Mt) tail_role(p, smem, cta - n_layer);
    else sampler_role(p, smem);
}

// create_upsample stage (wavenet/model.py:102-111): one conv2d_transpose(kernel (F,2), strides (F,1), 'same')
//   out[i*F + a][w] = in[i][w]*K[a][0] (+) fma(in[i][w-1], K[a][1])
extern "C" __global__ void wn_upsample_stage_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                    const float *__restrict__ K, long long rows_in, int F, int C)
{
    const long long total = rows_in * F * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int w = (int)(idx % C);
        long long ro = idx / C;
        int a = (int)(ro % F);
        long long i = ro / F;
        float x0 = in[i * C + w];
        float x1 = (w > 0) ? in[i * C + w - 1] : 0.0f;
        float v = fmul(x0, __ldg(K + a * 2 + 0));
        v = ffma(x1, __ldg(K + a * 2 + 1), v);
        out[idx] = v;
    }
}

// wavenet/ops.py:22-33
extern "C" __global__ void wn_mu_law_encode_kernel(const float *__restrict__ audio, long long n, float mu, int32_t *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float a = audio[i];
        float safe = fminf(fabsf(a), 1.0f);
        float mag = __fdiv_rn(log1pf(__fmul_rn(mu, safe)), log1pf(mu));
        float sgn = (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : 0.0f);
        float sig = __fmul_rn(sgn, mag);
        float v = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(sig, 1.0f), 2.0f), mu), 0.5f);
        out[i] = (int32_t)v;
    }
}

// wavenet/ops.py:36-47
extern "C" __global__ void wn_mu_law_decode_kernel(const float *__restrict__ in, long long n, float mu, int quantization,
                                                   float *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float sig = quantization ? __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(in[i], mu)), 1.0f) : in[i];
        float mag = __fmul_rn(__fdiv_rn(1.0f, mu), __fsub_rn(powf(__fadd_rn(1.0f, mu), fabsf(sig)), 1.0f));
        float sgn = (sig > 0.0f) ? 1.0f : ((sig < 0.0f) ? -1.0f : 0.0f);
        out[i] = __fmul_rn(sgn, mag);
    }
}
