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

extern __shared__ __align__(128) float g_smem[];

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
__device__ __forceinline__ void ll_store(u64 *p, float val, unsigned seq)
{
    u64 v = ((u64)seq << 32) | (u64)__float_as_uint(val);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Mailbox context: logical word address -> the grain copy on one die (wn_params.h)
struct MBox {
    u64 *const *tab0;
    u64 *const *tab1;
    u64 *const *mine;     // table of the reader's own die
    bool dual;
    __device__ __forceinline__ static u64 *at(u64 *const *tab, const u64 *logical)
    {
        const size_t w = reinterpret_cast<size_t>(logical) >> 3;
        return reinterpret_cast<u64 *>(__ldg(reinterpret_cast<const unsigned long long *>(tab) + (w >> 8))) + (w & 255);
    }
    __device__ __forceinline__ u64 *rd(const u64 *logical) const { return at(mine, logical); }
};
// resolved destination(s) of one posted word; resolve early (before the data exists), store late
struct MDst {
    u64 *p0, *p1;
};
__device__ __forceinline__ MDst mb_dst(const MBox &mb, const u64 *logical)
{
    MDst d;
    d.p0 = MBox::at(mb.tab0, logical);
    d.p1 = mb.dual ? MBox::at(mb.tab1, logical) : nullptr;
    return d;
}
__device__ __forceinline__ void ll_post(const MDst &d, float val, unsigned seq)
{
    ll_store(d.p0, val, seq);
    if (d.p1) ll_store(d.p1, val, seq);
}
__device__ __forceinline__ void ll_post(const MBox &mb, const u64 *logical, float val, unsigned seq)
{
    ll_post(mb_dst(mb, logical), val, seq);
}
__device__ __forceinline__ MBox make_mbox(const WnParams &p)
{
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    MBox mb;
    mb.tab0 = p.mb_tab[0];
    mb.tab1 = p.mb_tab[1];
    const int side = (p.sm_die != nullptr) ? (int)p.sm_die[smid] : 0;
    mb.mine = side ? p.mb_tab[1] : p.mb_tab[0];
    mb.dual = p.mb_dual != 0;
    return mb;
}
__device__ __forceinline__ int ld_volatile_i32(const int *p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// optional phase profile: per-phase clock64() deltas of thread 0, accumulated in REGISTERS (a global
// read-modify-write per mark would stall the warp for an L2 round trip and distort what it measures)
// and written to prof[cta*16 + phase] once, when the role returns normally.
struct Prof {
    long long *slot;
    long long last;
    long long acc[12];
    __device__ __forceinline__ Prof(long long *s) : slot(s), last(0) { for (int i = 0; i < 12; ++i) acc[i] = 0; }
    __device__ __forceinline__ void start() { if (slot) last = clock64(); }
    __device__ __forceinline__ void mark(int phase)
    {
        if (slot) { long long now = clock64(); acc[phase] += now - last; last = now; }
    }
    // accumulate a GPU-global nanosecond timestamp (low 40 bits): the difference of two CTAs' sums over the same steps
    // is the mean latency between the two events (cross-SM, unlike clock64)
    __device__ __forceinline__ void stamp(int idx)
    {
        if (slot) {
            unsigned long long g;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
            acc[idx] += (long long)(g & 0xffffffffffULL);
        }
    }
    __device__ __forceinline__ void flush(bool writer = (threadIdx.x == 0))
    {
        if (slot && writer)
            for (int i = 0; i < 12; ++i) slot[i] = acc[i];
    }
};

// prefetch helpers: volatile asm keeps these loads (and, through pin(), what is computed from them)
// ahead of the volatile polling loads, so their latency overlaps the wait instead of following it
__device__ __forceinline__ float ld_nc_f32(const float *p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_nc_f64(const double *p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void pin(float &v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void pin(double &v) { asm volatile("" : "+d"(v)); }

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

__device__ __forceinline__ float ll_wait(const MBox &mb, const u64 *logical, unsigned seq, Abort &ab)
{
    const u64 *p = mb.rd(logical);
    u64 v = ld_relaxed_u64(p);
    unsigned spins = 0;
    long long t0 = 0;
    while ((unsigned)(v >> 32) != seq) {
        if (((++spins) & 0x3ffu) == 0 && watchdog_check(ab, t0)) break;
        v = ld_relaxed_u64(p);
    }
    return __uint_as_float((unsigned)v);
}

// wait for n (<=4) words p[i*stride]; the loads are issued together so their latencies overlap.
// One poll = one L2 round trip (~350 cycles, profiles/r01_hop_latency.md).  Measured alternatives that were
// WORSE on the real kernel: two staggered polls per word in flight (+8 % step time), a delay between polling
// rounds (+6..18 %), a lazy single-thread "heads-up" watch on the previous stage (+10 %).
__device__ __forceinline__ void ll_wait_n(const MBox &mb, const u64 *logical, size_t stride, int n, unsigned seq, Abort &ab, float *out)
{
    const u64 *pp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pp[i] = (i < n) ? mb.rd(logical + i * stride) : nullptr;
    u64 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (i < n) ? ld_relaxed_u64(pp[i]) : ((u64)seq << 32);
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
            if (i < n && (unsigned)(v[i] >> 32) != seq) v[i] = ld_relaxed_u64(pp[i]);
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
// Thread-major matvec (layout: wn_params.h).  Per-thread view of one packed matrix, built once
// before the step loop so no index arithmetic (integer divisions!) sits on the sample chain.
struct MatT {
    const float *ws;     // shared-memory copy:  packed base + tid*V   (valid when smem)
    const float *wg;     // global image copy:   packed base + tid*V
    const float *xc;     // this thread's chunk of the padded input vector (shared memory)
    int grp;             // tid / tpc : column within a pass
    int n4, ch, u, tpc, npass, gpp, ncols, V;
    bool lead;           // tid % tpc == 0 : holds the result after the butterfly
    bool smem;
};

__device__ __forceinline__ int xpad(const WnMat &m, int k) { return (k / m.ch) * m.xstride + (k % m.ch); }

__device__ __forceinline__ MatT make_matt(const WnMat &m, const float *gimg, const float *xs)
{
    const int tid = threadIdx.x;
    MatT t;
    t.ws = g_smem + m.off + tid * m.V;
    t.wg = gimg + m.off + tid * m.V;
    t.xc = xs + (tid % m.t) * m.xstride;
    t.grp = tid / m.t;
    t.n4 = m.ch >> 2; t.ch = m.ch; t.u = m.u; t.tpc = m.t; t.npass = m.npass; t.gpp = m.gpp; t.ncols = m.ncols; t.V = m.V;
    t.lead = (tid % m.t) == 0;
    t.smem = m.in_smem != 0;
    return t;
}

// N4 float4 loads per thread, U independent fma sub-chains of N4/U float4 each (canonical chunks
// chunk*U .. chunk*U+U-1), combined by the first log2(U) levels of the ascending butterfly.
template <int N4, int U>
__device__ __forceinline__ float dot_regs(const float4 *__restrict__ w4, const float4 *__restrict__ x4)
{
    float4 wv[N4], xv[N4];
#pragma unroll
    for (int i = 0; i < N4; ++i) { wv[i] = w4[i * WN_NT]; xv[i] = x4[i]; }
    constexpr int PER = N4 / U;
    float acc[U];
#pragma unroll
    for (int s = 0; s < U; ++s) {
        float a = 0.0f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const float4 ww = wv[s * PER + i], xx = xv[s * PER + i];
            a = ffma(ww.x, xx.x, a);
            a = ffma(ww.y, xx.y, a);
            a = ffma(ww.z, xx.z, a);
            a = ffma(ww.w, xx.w, a);
        }
        acc[s] = a;
    }
#pragma unroll
    for (int off = 1; off < U; off <<= 1)
#pragma unroll
        for (int c = 0; c < U; c += 2 * off) acc[c] = fadd(acc[c], acc[c + off]);
    return acc[0];
}

// sub-chain count the device code implements for a chunk of n4 float4 (host mirrors this: wn_api.cu make_mat)
__host__ __device__ constexpr int wn_u_for_n4(int n4)
{
    return n4 == 2 ? 2 : n4 == 4 ? 4 : n4 == 8 ? 8 : n4 == 16 ? 8 : 1;
}

__device__ __forceinline__ float dot_thread(const MatT &m, const float *__restrict__ w)
{
    if (m.V == 4) {
        const float4 *w4 = reinterpret_cast<const float4 *>(w);
        const float4 *x4 = reinterpret_cast<const float4 *>(m.xc);
        switch (m.n4) {
        case 1: return dot_regs<1, 1>(w4, x4);
        case 2: return dot_regs<2, 2>(w4, x4);
        case 4: return dot_regs<4, 4>(w4, x4);
        case 8: return dot_regs<8, 8>(w4, x4);
        case 16: return dot_regs<16, 8>(w4, x4);
        case 5: return dot_regs<5, 1>(w4, x4);
        default: {
            float acc = 0.0f;
            for (int i = 0; i < m.n4; ++i) {
                const float4 ww = w4[i * WN_NT], xx = x4[i];
                acc = ffma(ww.x, xx.x, acc);
                acc = ffma(ww.y, xx.y, acc);
                acc = ffma(ww.z, xx.z, acc);
                acc = ffma(ww.w, xx.w, acc);
            }
            return acc;
        }
        }
    }
    float acc = 0.0f;
    for (int i = 0; i < m.ch; ++i) acc = ffma(w[i * WN_NT], m.xc[i], acc);
    return acc;
}

// All WN_NT threads must call it (warp shuffles inside).  epi(col, dot) runs on the lead lane of each column.
template <class F>
__device__ __forceinline__ void matvec(const MatT &m, F &&epi)
{
    for (int pass = 0; pass < m.npass; ++pass) {
        float acc = m.smem ? dot_thread(m, m.ws + (size_t)pass * m.ch * WN_NT) : dot_thread(m, m.wg + (size_t)pass * m.ch * WN_NT);
        for (int off = 1; off < m.tpc; off <<= 1) acc = fadd(acc, __shfl_xor_sync(FULL, acc, off));
        const int col = pass * m.gpp + m.grp;
        if (m.lead && col < m.ncols) epi(col, acc);
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
__device__ void layer_role(const WnParams &p, int l, int m)
{
    float *smem = g_smem;
    const int tid = threadIdx.x;
    const int N = p.N, L = p.L, R = p.R, M = p.M, Dm = p.Dm, Sm = p.Sm;
    const int cta = l * M + m;
    const float *gimg = p.layer_img + (size_t)cta * p.layer_img_floats;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.layer_smem_floats, &bar);

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
    const MBox mb = make_mbox(p);
    Prof pf(p.prof ? p.prof + (size_t)cta * 16 : nullptr);

    // ---- per-thread views and indices, computed once ------------------------------------------------
    const MatT mt_cur = make_matt(p.cur, gimg, xs_cur), mt_old = make_matt(p.old, gimg, xs_old);
    const MatT mt_lc = make_matt(p.lc, gimg, lcs), mt_gc = make_matt(p.gc, gimg, gvec);
    const MatT mt_dense = make_matt(p.dense, gimg, zs_dense), mt_skip = make_matt(p.skip, gimg, zs_skip);
    const int xp_cur = (tid < R) ? xpad(p.cur, tid) : 0;
    const int xp_old = (tid < R) ? xpad(p.old, tid) : 0;
    const int xp_lc = (p.C && tid < p.C) ? xpad(p.lc, tid) : 0;
    const int xp_zskip_gather = (tid < p.D) ? xpad(p.skip, tid) : 0;
    const int g_mm = (tid < p.D) ? tid / Dm : 0, g_j = (tid < p.D) ? tid % Dm : 0;   // z gather source
    // fg epilogue lane: column grp = 2*j + gate
    const int fg_grp = mt_cur.grp;
    const bool fg_valid = fg_grp < ncol2;
    const bool fg_gate = (fg_grp & 1) != 0;
    const bool fg_store = fg_valid && mt_cur.lead && !fg_gate;
    const int fg_j = fg_grp >> 1;
    const int xp_zd = fg_store ? xpad(p.dense, fg_j) : 0;
    const int xp_zs = fg_store ? xpad(p.skip, m * Dm + fg_j) : 0;
    const size_t rowx = (size_t)L * M * R, rowz = (size_t)L * M * Dm, rowa = (size_t)L * M * Sm;
    const u64 *mbx_in = p.mb_x + ((size_t)l * M) * R + tid;
    u64 *mbx_out = p.mb_x + ((size_t)(l + 1) * M + m) * R;
    u64 *mbz_out = p.mb_z + ((size_t)l * M + m) * Dm + fg_j;
    const u64 *mbz_in = p.mb_z + ((size_t)l * M + g_mm) * Dm + g_j;
    const u64 *mba_in = p.mb_acc + ((size_t)(l > 0 ? l - 1 : 0) * M + m) * Sm;
    u64 *mba_out = p.mb_acc + ((size_t)l * M + m) * Sm;

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
            xs_old[xp_old] = v;
        }
        if (p.C && tid < p.C) {
            long idx = (long)tn - 1 - p.lc_shift;
            float v = 0.0f;
            if (p.lc_up != nullptr && idx >= 0 && idx < p.t_lc) v = __ldg(p.lc_up + ((size_t)b * p.t_lc + idx) * p.C + tid);
            lcs[xp_lc] = v;
        }
        __syncthreads();
        float *pre_b = pre + b * ncol2;
        const float *bias_b = bfgN + b * ncol2;
        matvec(mt_old, [&](int col, float dot) { pre_b[col] = fadd(bias_b[col], dot); });
        if (p.C) {
            __syncthreads();
            matvec(mt_lc, [&](int col, float dot) { pre_b[col] = fadd(pre_b[col], dot); });
        }
        __syncthreads();
    };

    // ---- prologue: fold the speaker embedding into the biases (model.py:71-73,181-212), pre for t = 0
    for (int b = 0; b < N; ++b) {
        if (p.G) {
            if (tid < p.G) gvec[xpad(p.gc, tid)] = __ldg(p.gc_table + (size_t)p.gc_id[b] * p.G + tid);
            __syncthreads();
            matvec(mt_gc, [&](int col, float dot) { bfgN[b * ncol2 + col] = fadd(bfg[col], dot); });
        } else {
            if (tid < ncol2) bfgN[b * ncol2 + tid] = bfg[tid];
        }
        __syncthreads();
        compute_pre(b, 0);
    }

    // ---- main loop
    for (int t = 0; t < p.T; ++t) {
        const unsigned seq = (unsigned)t + 1u;
        for (int b = 0; b < N; ++b) {
            if (t >= p.T_row[b]) continue;
            pf.start();
            // 1. wait for the layer input: sum of the partial residual outputs of layer l-1
            if (tid < R) {
                float q[4];
                ll_wait_n(mb, mbx_in + b * rowx, (size_t)R, nin, seq, ab, q);
                float v = q[0];
                for (int i = 1; i < nin; ++i) v = fadd(v, q[i]);
                xs_cur[xp_cur] = v;
                xraw[tid] = v;
            }
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(0);

            // 2. filter/gate for the current tap + gated activation (model.py:68-69,86)
            {
                float acc = mt_cur.smem ? dot_thread(mt_cur, mt_cur.ws) : dot_thread(mt_cur, mt_cur.wg);
                for (int off = 1; off < mt_cur.tpc; off <<= 1) acc = fadd(acc, __shfl_xor_sync(FULL, acc, off));
                float pv = fg_valid ? pre[b * ncol2 + fg_grp] : 0.0f;
                float fg = fadd(pv, acc);
                float a = act_fg(fg, fg_gate);
                float other = __shfl_xor_sync(FULL, a, mt_cur.tpc);   // partner column (filter <-> gate)
                if (fg_store) {
                    float z = fmul(a, other);
                    zs_dense[xp_zd] = z;
                    zs_skip[xp_zs] = z;
                    if (M > 1) ll_post(mb, mbz_out + b * rowz, z, seq);
                }
            }
            __syncthreads();
            pf.mark(1);

            // 3. partial dense 1x1 + residual (model.py:89,98-101) -> mailbox of layer l+1
            if (l + 1 < L) {
                u64 *dst = mbx_out + b * rowx;
                matvec(mt_dense, [&](int r, float dot) {
                    float v = (m == 0) ? fadd(fadd(xraw[r], bd[r]), dot) : dot;
                    ll_post(mb, dst + r, v, seq);
                });
            }
            pf.mark(2);
            // ---- everything below is off the sample-to-sample critical chain ----
            // 4. push x_l(t) into the private dilation-queue ring (model.py:145)
            if (d >= 2 && tid < R) __stcg(ring_cta + ((size_t)b * d + (t % d)) * R + tid, xraw[tid]);
            // 5. gather the sibling CTAs' gated activations
            if (M > 1 && tid < p.D && g_mm != m) zs_skip[xp_zskip_gather] = ll_wait(mb, mbz_in + b * rowz, seq, ab);
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(3);
            // 6. skip 1x1 (model.py:94-96) + running sum over layers (model.py:157)
            {
                const u64 *src = mba_in + b * rowa;
                u64 *dst = mba_out + b * rowa;
                matvec(mt_skip, [&](int c, float dot) {
                    float v = fadd(bs[c], dot);
                    if (l > 0) v = fadd(ll_wait(mb, src + c, seq, ab), v);
                    ll_post(mb, dst + c, v, seq);
                });
            }
            pf.mark(4);
            // 7. pre-activations of the next step
            if (t + 1 < p.T_row[b]) compute_pre(b, t + 1);
            else __syncthreads();
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(5);
        }
    }
    pf.flush();
}

// =============================================================================================
// Tail CTA: relu -> conv1 (S->S) -> relu -> partial conv2 (model.py:158-165)
__device__ void tail_role(const WnParams &p, int mt)
{
    float *smem = g_smem;
    const int tid = threadIdx.x;
    const int N = p.N, L = p.L, M = p.M, Sm = p.Sm, S = p.S;
    const float *gimg = p.tail_img + (size_t)mt * p.tail_img_floats;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.tail_smem_floats, &bar);
    const float *b1 = smem + p.off_b1;
    float *sc = smem + p.tail_smem_floats;
    float *as1 = sc + p.ts.as1, *c1s = sc + p.ts.c1s;
    for (int i = tid; i < p.post1.xlen; i += WN_NT) as1[i] = 0.0f;
    for (int i = tid; i < p.post2.xlen; i += WN_NT) c1s[i] = 0.0f;
    __syncthreads();
    const MatT mt_p1 = make_matt(p.post1, gimg, as1), mt_p2 = make_matt(p.post2, gimg, c1s);
    // this thread polls acc words c = tid, tid + NT, ... (at most 4 supported per thread, else loop)
    const size_t rowa = (size_t)L * M * Sm;
    const u64 *src0 = p.mb_acc + ((size_t)(L - 1) * M) * Sm;   // [M][Sm] == S contiguous words
    u64 *dst0 = p.mb_c2 + (size_t)mt * p.O;
    const int xp1_lead = mt_p1.lead ? xpad(p.post2, mt_p1.grp < p.St ? mt_p1.grp : 0) : 0;   // single pass: col == grp
    Abort ab{p.status, 0};
    const MBox mb = make_mbox(p);
    Prof pf(p.prof ? p.prof + (size_t)(p.L * p.M + mt) * 16 : nullptr);
    for (int t = 0; t < p.T; ++t) {
        const unsigned seq = (unsigned)t + 1u;
        for (int b = 0; b < N; ++b) {
            if (t >= p.T_row[b]) continue;
            pf.start();
            const u64 *src = src0 + b * rowa;
            for (int c = tid; c < S; c += WN_NT) as1[xpad(p.post1, c)] = relu32(ll_wait(mb, src + c, seq, ab));
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(0);
            if (mt_p1.npass == 1)
                matvec(mt_p1, [&](int c, float dot) { c1s[xp1_lead] = relu32(fadd(b1[c], dot)); });
            else
                matvec(mt_p1, [&](int c, float dot) { c1s[xpad(p.post2, c)] = relu32(fadd(b1[c], dot)); });
            __syncthreads();
            pf.mark(1);
            u64 *dst = dst0 + (size_t)b * p.Mt * p.O;
            matvec(mt_p2, [&](int o, float dot) { ll_post(mb, dst + o, dot, seq); });
            __syncthreads();
            pf.mark(2);
        }
    }
    pf.flush();
}

// float64 softmax (model.py:243) -> fp32; temperature + categorical draw (generate.py:219-231).
// Called by every thread of the sampler CTA; returns the drawn id (as float) to all of them.
__device__ float mulaw_draw_cta(const float *c2s, int Q, float temperature, double u64v, float *misc, double *red, double *cdf,
                                float *probs_out = nullptr)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    if (warp == 0) {      // levels 32, 64, 128 of the butterfly: lanes 0..nw-1 hold the warp sums
        double v = (lane < nw) ? red[lane] : 0.0;
        for (int off = 1; off < nw; off <<= 1) v = __dadd_rn(v, __shfl_xor_sync(FULL, v, off));
        if (lane == 0) red[8] = v;
    }
    __syncthreads();
    const double den = red[8];
    float pr = act ? (float)__ddiv_rn(e, den) : 0.0f;
    if (probs_out != nullptr && act) probs_out[tid] = pr;          // predict_proba_incremental's return value (model.py:243)
    float s = fdiv(wn::log32(pr), temperature);
    float a = s;
    for (int off = 1; off < 32; off <<= 1) {
        float o = __shfl_xor_sync(FULL, a, off);
        a = (lane & off) ? wn::logaddexp32(o, a) : wn::logaddexp32(a, o);
    }
    if (lane == 0 && act) misc[16 + warp] = a;
    __syncthreads();
    if (warp == 0) {
        float v = (lane < nw) ? misc[16 + lane] : 0.0f;
        for (int off = 1; off < nw; off <<= 1) {
            float o = __shfl_xor_sync(FULL, v, off);
            v = (lane & off) ? wn::logaddexp32(o, v) : wn::logaddexp32(v, o);
        }
        if (lane == 0) misc[1] = v;
    }
    __syncthreads();
    const float lse = misc[1];
    // cumulative sum of float64(q), pinned order: Kogge-Stone scan inside each block of 32
    // (v_j += v_{j-off}, off = 1..16), block totals added left to right as the block's base.
    double v = act ? (double)wn::exp32(fsub(s, lse)) : 0.0;
    for (int off = 1; off < 32; off <<= 1) {
        double o = __shfl_up_sync(FULL, v, off);
        if (lane >= off) v = __dadd_rn(v, o);
    }
    if (lane == 31 && act) red[warp] = v;
    __syncthreads();
    if (warp > 0) {
        double base = red[0];
        for (int w = 1; w < warp; ++w) base = __dadd_rn(base, red[w]);
        v = __dadd_rn(base, v);
    }
    if (tid == Q - 1) cdf[0] = v;
    __syncthreads();
    const double total = cdf[0];
    int pred = act && (__ddiv_rn(v, total) <= u64v);      // searchsorted(side='right')
    int cnt = __syncthreads_count(pred);
    if (cnt > Q - 1) cnt = Q - 1;
    return (float)cnt;
}

// =============================================================================================
// Sampler CTA
__device__ void sampler_role(const WnParams &p)
{
    float *smem = g_smem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, R = p.R, O = p.O, Q = p.Q, nr = p.nr_mix, ifw = p.ifw;
    const float *gimg = p.samp_img;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.samp_smem_floats, &bar);
    const float *b2 = smem + p.off_b2;
    float *sc = smem + p.samp_smem_floats;
    float *c2s = sc + p.ss.c2s, *cq = sc + p.ss.cq, *cqx = sc + p.ss.cqx, *qs = sc + p.ss.qs;
    int *ids = reinterpret_cast<int *>(sc + p.ss.ids);
    double *cdf = reinterpret_cast<double *>(sc + p.ss.cdf);
    double *red = reinterpret_cast<double *>(sc + p.ss.red);     // 16 doubles
    float *misc = sc + p.ss.misc;                                 // 32 floats
    Abort ab{p.status, 0};
    const MBox mb = make_mbox(p);
    Prof pf(p.prof ? p.prof + (size_t)(p.L * p.M + p.Mt) * 16 : nullptr);

    for (int i = tid; i < N * ifw; i += WN_NT) cq[i] = 0.0f;
    for (int i = tid; i < p.causal.xlen; i += WN_NT) cqx[i] = 0.0f;
    for (int i = tid; i < 2 * N; i += WN_NT) ids[i] = -1;
    __syncthreads();
    const MatT mt_c = make_matt(p.causal, gimg, cqx);
    const int xp_cq = (p.scalar_input && tid < ifw) ? xpad(p.causal, tid) : 0;
    const size_t rowx = (size_t)p.L * p.M * R;

    // push x_in into row b's causal queue, run the causal conv, post to layer 0 with tag seq
    auto feed = [&](int b, float x_in, unsigned seq) {
        u64 *dst = p.mb_x + b * rowx;
        if (p.scalar_input) {
            float v = 0.0f;
            if (tid < ifw) v = (tid < ifw - 1) ? cq[b * ifw + tid + 1] : x_in;
            __syncthreads();
            if (tid < ifw) { cq[b * ifw + tid] = v; cqx[xp_cq] = v; }
            __syncthreads();
            matvec(mt_c, [&](int r, float dot) { ll_post(mb, dst + r, dot, seq); });
        } else {
            int prev = ids[2 * b + 1];
            int cur = (int)x_in;
            __syncthreads();
            if (tid == 0) { ids[2 * b] = prev; ids[2 * b + 1] = cur; }
            if (tid < R) {
                float a = (prev >= 0) ? __ldg(p.wc_onehot + ((size_t)0 * Q + prev) * R + tid) : 0.0f;
                float bb = (cur >= 0 && cur < Q) ? __ldg(p.wc_onehot + ((size_t)1 * Q + cur) * R + tid) : 0.0f;
                ll_post(mb, dst + tid, fadd(a, bb), seq);
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
            pf.start();
            // prefetch this step's uniforms and the next forced input while the network runs
            float gum = 0.0f, logistic = 0.0f, next_forced = 0.0f;
            double u64v = 0.0;
            if (p.scalar_input) {
                const float *u = (const float *)p.uniforms + ((size_t)b * p.T + t) * (nr + 1);
                if (warp == 0 && lane < nr) gum = wn::log32(-wn::log32(ld_nc_f32(u + lane)));
                if (tid == 0) { float u2 = ld_nc_f32(u + nr); logistic = fsub(wn::log32(u2), wn::log32(fsub(1.0f, u2))); }
            } else {
                u64v = ld_nc_f64((const double *)p.uniforms + (size_t)b * p.T + t);
            }
            const bool has_next = (t + 1 < p.T_row[b]);
            if (has_next && t + 1 < p.n_forced) next_forced = ld_nc_f32(p.forced + (size_t)b * p.n_forced + t + 1);
            pin(gum); pin(logistic); pin(next_forced); pin(u64v);

            // conv2 output: bias + the Mt partial sums in order
            if (tid < O) {
                float v = b2[tid];
                const u64 *src = p.mb_c2 + ((size_t)b * p.Mt) * O + tid;
                for (int m0 = 0; m0 < p.Mt; m0 += 4) {
                    float q[4];
                    int n = (p.Mt - m0 < 4) ? (p.Mt - m0) : 4;
                    ll_wait_n(mb, src + (size_t)m0 * O, (size_t)O, n, seq, ab, q);
                    for (int i = 0; i < n; ++i) v = fadd(v, q[i]);
                }
                c2s[tid] = v;
                if (p.out_logits) p.out_logits[((size_t)b * p.T + t) * O + tid] = v;
            }
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(0);

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
                sample = mulaw_draw_cta(c2s, Q, p.temperature, u64v, misc, red, cdf);
            }
            pf.mark(1);
            if (tid == 0) p.out_samples[(size_t)b * p.T + t] = sample;
            if (has_next) feed(b, (t + 1 < p.n_forced) ? next_forced : sample, seq + 1u);
            else __syncthreads();
            pf.mark(2);
        }
    }
    pf.flush();
}

// wavenet/mixture.py:84-114 sample_from_discretized_mix_logistic on a tensor of logits (rows, 3*nr) with uniforms (rows, nr + 1):
// the same pinned arithmetic as the in-kernel draw (Gumbel-max over y - log(-log u), first maximum wins like tf.argmax, the
// selected mean / clamped log-scale, mean + exp(log_scale) * (log u - log(1 - u)), clip to [-1, 1]).  One thread per row.
extern "C" __global__ void wn_mol_sample_kernel(const float *__restrict__ y, const float *__restrict__ u, long long rows, int nr,
                                                float log_scale_min, float *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const float *yr = y + i * (3 * nr), *ur = u + i * (nr + 1);
        int best = 0;
        float bestv = 0.0f;
        for (int k = 0; k < nr; ++k) {
            const float g = fsub(yr[k], wn::log32(-wn::log32(ur[k])));
            if (k == 0 || g > bestv) { bestv = g; best = k; }
        }
        const float mean = yr[nr + best];
        float ls = yr[2 * nr + best];
        if (!(ls > log_scale_min)) ls = log_scale_min;
        const float u2 = ur[nr];
        const float d = fsub(wn::log32(u2), wn::log32(fsub(1.0f, u2)));
        float x = fadd(mean, fmul(wn::exp32(ls), d));
        x = fmaxf(x, -1.0f);
        x = fminf(x, 1.0f);
        out[i] = x;
    }
}

// wavenet/mixture.py:27-81 discretized_mix_logistic_loss, forward only, on a tensor of network outputs y_hat (rows, 3*nr) and
// targets y (rows): per-row loss -logsumexp_k(log_prob_k + log_softmax(logit)_k) with the reference's three tf.where branches
// (y < -0.999: log cdf_plus; y > 0.999: log(1 - cdf_min); else log(max(cdf_delta, 1e-12)) where cdf_delta > 1e-5, the mid-point
// log-pdf otherwise).  fp32 like the reference's graph; loss_out (rows) and / or sum_out (one double, atomically accumulated).
extern "C" __global__ void wn_mol_loss_kernel(const float *__restrict__ y_hat, const float *__restrict__ y, long long rows, int nr,
                                              float log_scale_min, float half_bin, float log_half_classes,
                                              float *__restrict__ loss_out, double *__restrict__ sum_out)
{
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const float *yr = y_hat + i * (3 * nr);
        const float t = y[i];
        float mx = -INFINITY;
        for (int k = 0; k < nr; ++k) mx = fmaxf(mx, yr[k]);
        float se = 0.0f;
        for (int k = 0; k < nr; ++k) se += expf(yr[k] - mx);
        const float lse_logits = mx + logf(se);
        float amax = -INFINITY, asum = 0.0f;              // streaming logsumexp over the components
        for (int k = 0; k < nr; ++k) {
            const float ls = fmaxf(yr[2 * nr + k], log_scale_min);
            const float c = t - yr[nr + k], inv = expf(-ls);
            const float pin = inv * (c + half_bin), min_ = inv * (c - half_bin), mid = inv * c;
            float logp;
            if (t < -0.999f) {
                logp = pin - (fmaxf(pin, 0.0f) + log1pf(expf(-fabsf(pin))));
            } else if (t > 0.999f) {
                logp = -(fmaxf(min_, 0.0f) + log1pf(expf(-fabsf(min_))));
            } else {
                const float delta = 1.0f / (1.0f + expf(-pin)) - 1.0f / (1.0f + expf(-min_));
                if (delta > 1e-5f) logp = logf(fmaxf(delta, 1e-12f));
                else logp = mid - ls - 2.0f * (fmaxf(mid, 0.0f) + log1pf(expf(-fabsf(mid)))) - log_half_classes;
            }
            const float a = logp + (yr[k] - lse_logits);
            if (a > amax) { asum = asum * expf(amax - a) + 1.0f; amax = a; }
            else asum += expf(a - amax);
        }
        const float loss = -(amax + logf(asum));
        if (loss_out) loss_out[i] = loss;
        acc += (double)loss;
    }
    if (sum_out) {
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if ((threadIdx.x & 31) == 0) atomicAdd(sum_out, acc);
    }
}

#include "wn_kernel_static.cuh"
#include "wn_kernel_ws.cuh"
#include "wn_kernel_v2.cuh"
#include "wn_step.cuh"

// LL-mailbox ping-pong between CTA 0 and CTA 1: average round trip in clock cycles (diagnostic).
__device__ void pingpong_role(u64 *box, int iters, long long *out)
{
    const int me = blockIdx.x;
    if (threadIdx.x != 0 || me > 1) return;
    u64 *mine = box + me * 16, *theirs = box + (1 - me) * 16;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            ll_store(theirs, 1.0f, (unsigned)i);
            while ((unsigned)(ld_relaxed_u64(mine) >> 32) != (unsigned)i) {}
        } else {
            while ((unsigned)(ld_relaxed_u64(mine) >> 32) != (unsigned)i) {}
            ll_store(theirs, 1.0f, (unsigned)i);
        }
    }
    if (me == 0) out[0] = (clock64() - t0) / iters;
}

}  // namespace

extern "C" __global__ void wn_pingpong_kernel(unsigned long long *box, int iters, long long *out)
{
    pingpong_role(box, iters, out);
}

// Diagnostic: CTA 0 ping-pongs with every other CTA in turn; out[k] = round trip cycles with CTA k,
// out[grid + k] = SM id of CTA k.  mode 0: ld.relaxed.gpu / st.relaxed.gpu, 1: ld.volatile / st.volatile,
// 2: ld.global.cg / st.global.cg, 3: relaxed with 4 staggered poller lanes.
extern "C" __global__ void wn_pingpong_all_kernel(unsigned long long *box, int iters, long long *out, int mode)
{
    const int me = blockIdx.x, G = gridDim.x;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) out[G + me] = smid;
    auto ld = [&](const u64 *p) -> u64 {
        u64 v;
        if (mode == 1) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        else if (mode == 2) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    };
    auto st = [&](u64 *p, u64 v) {
        if (mode == 1) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
        else if (mode == 2) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
        else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    };
    // mailbox pair for partner k: box[(2k)*32] (to CTA k) and box[(2k+1)*32] (to CTA 0), 256 B apart
    const int lane = threadIdx.x;
    if (me == 0) {
        for (int k = 1; k < G; ++k) {
            u64 *to = box + (size_t)(2 * k) * 32, *from = box + (size_t)(2 * k + 1) * 32;
            long long t0 = clock64();
            for (int i = 1; i <= iters; ++i) {
                if (lane == 0) st(to, (u64)i);
                if (mode == 3) {
                    // lanes 0..3 poll the same word, started a quarter period apart
                    bool seen = false;
                    if (lane < 4) { for (int w = 0; w < lane * 8; ++w) __nanosleep(0); }
                    while (true) {
                        if (lane < 4 && !seen) seen = (ld(from) == (u64)i);
                        if (__any_sync(0xffffffffu, seen)) break;
                    }
                } else if (lane == 0) {
                    while (ld(from) != (u64)i) {}
                }
                __syncwarp();
            }
            if (lane == 0) out[k] = (clock64() - t0) / iters;
        }
    } else {
        u64 *mine = box + (size_t)(2 * me) * 32, *back = box + (size_t)(2 * me + 1) * 32;
        if (lane == 0)
            for (int i = 1; i <= iters; ++i) {
                while (ld(mine) != (u64)i) {}
                st(back, (u64)i);
            }
    }
}

// =============================================================================================
// Role of this CTA.  Roles are numbered along the sample chain (layers, tail, sampler).  With a die map,
// CTAs running on die 0 claim roles from the front of the list and CTAs on die 1 from the back, so the chain
// crosses the die boundary twice per sample instead of on about every second hop (writers on the reader's
// die save ~160 cycles per hop, profiles/r01_hop_latency.md).  Mailboxes are addressed by role, so no CTA
// needs to know who claimed what.  status[4], status[5]: claim counters, zeroed before every launch.
__device__ __forceinline__ int claim_role(const WnParams &p)
{
    __shared__ int s_role;
    if (threadIdx.x == 0) {
        int role = (int)blockIdx.x;
        if (p.sm_die != nullptr) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            role = (p.sm_die[smid] == 0) ? atomicAdd(p.status + 4, 1) : p.grid - 1 - atomicAdd(p.status + 5, 1);
        }
        s_role = role;
    }
    __syncthreads();
    return s_role;
}

extern "C" __global__ void __launch_bounds__(WN_NT, 1) wn_persistent_kernel(const __grid_constant__ WnParams p)
{
    const int cta = claim_role(p);
    const int n_layer = p.L * p.M;
    if (cta < n_layer) layer_role(p, cta / p.M, cta % p.M);
    else if (cta < n_layer + p.Mt) tail_role(p, cta - n_layer);
    else sampler_role(p);
}

template <class SH>
__global__ void __launch_bounds__(WN_NT, 1) wn_persistent_kernel_s(const __grid_constant__ WnParams p)
{
    const int cta = claim_role(p);
    constexpr int n_layer_per = SH::M;
    const int n_layer = p.L * n_layer_per;
    if (cta < n_layer) {
        if constexpr (SH::WS) layer_role_ws<SH>(p, cta / n_layer_per, cta % n_layer_per);
        else layer_role_s<SH>(p, cta / n_layer_per, cta % n_layer_per);
    }
    else if (cta < n_layer + SH::Mt) tail_role_s<SH>(p, cta - n_layer);
    else sampler_role_s<SH>(p);
}

// Diagnostic: cycles per polling round.  `warps` warps x `lanes` lanes each keep K strong loads in flight
// (distinct 8-byte words, the kernel's own mailbox pattern) and wait for all of them, `iters` times.
extern "C" __global__ void wn_pollbench_kernel(const unsigned long long *box, int iters, int warps, int lanes, int K, long long *out)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp >= warps || lane >= lanes) return;
    const u64 *p = box + (size_t)blockIdx.x * 4096 + tid;
    u64 acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        u64 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (k < K) ? ld_relaxed_u64(p + k * 128 + (acc & 1)) : 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += v[k];
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = (t1 - t0) / iters + (acc == 0x1234567 ? 1 : 0);
}

// Diagnostic: CTA 0 ping-pongs with partner CTAs part[0..np-1]; for each partner all (X, Y) combinations of
// ng grains: partner's inbox in grain X, CTA 0's inbox in grain Y.  out[(pi*ng + X)*ng + Y] = round trip cycles.
extern "C" __global__ void wn_pingpong_grid_kernel(unsigned long long *box, const int *part, int np, int ng, int iters,
                                                   long long *out, unsigned *smids)
{
    const int me = blockIdx.x;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x != 0) return;
    smids[me] = smid;
    unsigned tag = 0;
    for (int pi = 0; pi < np; ++pi) {
        const int k = part[pi];
        if (me != 0 && me != k) continue;
        for (int X = 0; X < ng; ++X)
            for (int Y = 0; Y < ng; ++Y) {
                u64 *to_k = box + (size_t)X * 256 + 8 * (pi + 1);       // word inside grain X
                u64 *to_0 = box + (size_t)Y * 256 + 8 * (pi + 1) + 4;   // word inside grain Y
                long long t0 = clock64();
                for (int i = 1; i <= iters; ++i) {
                    tag = (unsigned)(((pi * ng + X) * ng + Y) * iters + i);     // identical on both sides
                    if (me == 0) {
                        ll_store(to_k, 1.0f, tag);
                        while ((unsigned)(ld_relaxed_u64(to_0) >> 32) != tag) {}
                    } else {
                        while ((unsigned)(ld_relaxed_u64(to_k) >> 32) != tag) {}
                        ll_store(to_0, 1.0f, tag);
                    }
                }
                if (me == 0) out[((size_t)pi * ng + X) * ng + Y] = (clock64() - t0) / iters;
            }
    }
}

// Die calibration by ping-pong.  One CTA per SM.  A job = {smid_a, smid_b, first grain, grain
// count, out offset}: the CTAs running on SMs a and b bounce an LL word through each grain (both inboxes in the
// same 2 KB grain) and SM a records the round trip in out[offset + g].  A round trip is ~900 cycles only when both
// SMs sit on the die whose L2 homes the grain, ~1300-2100 otherwise.  Every wait is bounded (abort flag).
struct WnCalibJob {
    int a, b, g0, ng, out0;
};
extern "C" __global__ void wn_calib_kernel(unsigned long long *grains, const WnCalibJob *jobs, int n_jobs, int iters,
                                           unsigned *out, int *abort_flag)
{
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x != 0) return;
    for (int j = 0; j < n_jobs; ++j) {
        const WnCalibJob jb = jobs[j];
        const bool is_a = (int)smid == jb.a, is_b = (int)smid == jb.b;
        if (!is_a && !is_b) continue;
        for (int g = 0; g < jb.ng; ++g) {
            u64 *base = grains + (size_t)(jb.g0 + g) * 256 + 2 * (j & 63);
            u64 *to_b = base, *to_a = base + 1;
            long long t0 = clock64();
            for (int i = 1; i <= iters + 2; ++i) {
                if (i == 3) t0 = clock64();                       // two warm-up exchanges
                const unsigned tag = ((unsigned)(j + 1) << 12) + (unsigned)(g * 64 + i);
                unsigned spins = 0;
                if (is_a) {
                    ll_store(to_b, 0.0f, tag);
                    while ((unsigned)(ld_relaxed_u64(to_a) >> 32) != tag)
                        if ((++spins & 0xfff) == 0 && (ld_volatile_i32(abort_flag) || spins > (1u << 22))) { *abort_flag = 1; return; }
                } else {
                    while ((unsigned)(ld_relaxed_u64(to_b) >> 32) != tag)
                        if ((++spins & 0xfff) == 0 && (ld_volatile_i32(abort_flag) || spins > (1u << 22))) { *abort_flag = 1; return; }
                    ll_store(to_a, 0.0f, tag);
                }
            }
            if (is_a) out[jb.out0 + g] = (unsigned)((clock64() - t0) / iters);
        }
    }
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
// Noise of the mixture-of-logistics draw (wavenet/mixture.py:96-98, 107-108) for every (row, step): u -> log(-log u) for the nr
// Gumbel-max values, u -> log u - log(1 - u) for the logistic.  Same pinned log32 as the in-kernel draw; computed once before the
// launch so that two dependent log32 chains per row-step leave the layer-0 helper group's loop (its slowest stage at 8 rows).
extern "C" __global__ void wn_noise_prep_kernel(const float *__restrict__ u, float *__restrict__ out, long long n, int nr)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % (nr + 1));
        const float v = ld_nc_f32(u + i);
        out[i] = (j < nr) ? wn::log32(-wn::log32(v)) : fsub(wn::log32(v), wn::log32(fsub(1.0f, v)));
    }
}

extern "C" __global__ void wn_mu_law_encode_kernel(const float *__restrict__ audio, long long n, float mu, int32_t *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float a = audio[i];
        float safe = fminf(fabsf(a), 1.0f);
        float mag = __fdiv_rn(wn::log1p32(__fmul_rn(mu, safe)), wn::log1p32(mu));      // pinned: integer codes are bit-exact vs the oracle
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
