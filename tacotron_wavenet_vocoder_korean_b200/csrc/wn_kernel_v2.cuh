// wn_kernel_v2.cuh -- round-2 layer kernel of the WaveNet sample loop: thread-block clusters + distributed shared memory.
//
// Why (profiles/r02_probe_cluster.md): a layer-to-layer hop through an L2 mailbox costs ~930 cycles in the live kernel,
// the same 4-senders-to-4-receivers exchange through DSMEM (st.async + mbarrier complete_tx) ~510 including the combine.
// B200 co-schedules at most 15 clusters of 8 CTAs (120 CTAs = 30 layers x M = 4; 16-CTA clusters: only 7), so
//   * kernel A (this file, `wn_layers_kernel_v2`): 120 CTAs, cluster = 2 consecutive layers x 4 CTAs, 384 threads:
//       CHAIN group (warps 8-11): wait input -> combine -> filter/gate from registers -> tanh*sigmoid -> partial dense from
//         registers -> send.  Even layer -> odd layer: DSMEM.  Odd layer -> next cluster: LL mailbox in L2 (wn_kernel.cu).
//       HELPER group (warps 0-7): dilation-ring push, skip 1x1 from REGISTERS + running skip sum, pre-activations of the
//         next step (dilated tap + local condition) from shared memory.  The siblings' gated activations arrive by DSMEM.
//   * kernel B (`wn_tail_kernel_v2`): the tail / sampler roles of wn_kernel_static.cuh on the 28 SMs the clusters cannot
//       use, launched on a second stream; it talks to kernel A through the same L2 mailboxes as before.
//
// The evaluation plan (DESIGN.md "Pinned arithmetic") is IDENTICAL to wn_persistent_kernel_s<ShapeCfg2>: every dot
// product is the same set of canonical 4-element fma chains combined by the same ascending butterfly, only the
// assignment of chunks to lanes changed:
//   filter/gate: lane = canonical chunk (k = 4*lane .. 4*lane+3), a warp owns 8 columns (4 channels x {filter, gate});
//     ONE conflict-free LDS.128 of x per thread instead of 8 broadcast ones (256 -> 32 LSU wavefronts per row-step),
//     partial sums are reduced with a transposing butterfly (9 shuffles, the first three levels halve the live values);
//   dense: lane & 7 = canonical chunk of the CTA's 32 gated channels, 4 outputs per lane.
// Included by wn_kernel.cu after wn_kernel_ws.cuh (uses its mbarrier helpers).
#pragma once
#include <type_traits>

// Register split between the groups (setmaxnreg, one warpgroup-aligned instruction per group): 256 x HELPER + 128 x CHAIN
// <= 65536.  The chain group's filter/gate phase wants its 16 LDS.128 of the layer input in flight over 96 weight registers;
// at 168 ptxas splits them into two dependent batches.  Measured (profiles/r02_setmaxnreg_sweep.md): 136/232 starves the
// helper (8 rows 24.6 us), 160/184 is best at every row count.
constexpr int V2_REGS_HELPER = 160, V2_REGS_CHAIN = 184;
constexpr int V2_NT = 384;       // threads per layer CTA: 8 helper warps + 4 chain warps (168 registers per thread at launch)
constexpr int V2_HALF = 256;     // helper threads (warps 0-7); the chain group is warps 8-11
constexpr int V2_CHAIN = 128;
constexpr int V2_CS = 8;         // CTAs per cluster: 2 layers x M = 4 (the portable shape; 16 = 4 layers where 7 of them fit)

__device__ __forceinline__ unsigned v2_cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned v2_cluster_id()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t v2_mapa(uint32_t addr, unsigned rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void v2_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void v2_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 4-byte store into a peer CTA's shared memory; completes `bytes` on the peer's mbarrier when it lands
__device__ __forceinline__ void v2_st_async(uint32_t raddr, float v, uint32_t rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(__float_as_uint(v)),
                 "r"(rbar)
                 : "memory");
}
// named barrier over the 256 threads of one group, OR-reducing the abort flag
__device__ __forceinline__ int v2_group_sync_or(int id, int flag)
{
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "barrier.red.or.pred q, %2, 256, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(flag), "r"(id)
        : "memory");
    return r;
}

// 8 partial column sums per lane, lane = canonical chunk: returns the full sum of column (lane & 7) in every lane.
// Level `off` of the ascending butterfly a[c] += a[c ^ off] (oracle mv_plan); at the first three levels each lane keeps
// the columns whose index bit equals its lane bit and ships the others, so 4 + 2 + 1 shuffles instead of 3 x 8.
__device__ __forceinline__ float v2_reduce8(const float (&v)[8], int lane)
{
    const bool b0 = (lane & 1) != 0, b1 = (lane & 2) != 0, b2 = (lane & 4) != 0;
    float n4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float keep = b0 ? v[2 * k + 1] : v[2 * k];
        const float send = b0 ? v[2 * k] : v[2 * k + 1];
        n4[k] = fadd(keep, __shfl_xor_sync(FULL, send, 1));
    }
    float n2[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float keep = b1 ? n4[2 * q + 1] : n4[2 * q];
        const float send = b1 ? n4[2 * q] : n4[2 * q + 1];
        n2[q] = fadd(keep, __shfl_xor_sync(FULL, send, 2));
    }
    const float keep = b2 ? n2[1] : n2[0];
    const float send = b2 ? n2[0] : n2[1];
    float r = fadd(keep, __shfl_xor_sync(FULL, send, 4));
    r = fadd(r, __shfl_xor_sync(FULL, r, 8));
    r = fadd(r, __shfl_xor_sync(FULL, r, 16));
    return r;
}
// 4 partial sums per lane over 8 lanes (lane & 7 = canonical chunk): full sum of value (lane & 3) in every lane
__device__ __forceinline__ float v2_reduce4(const float (&v)[4], int lane)
{
    const bool b0 = (lane & 1) != 0, b1 = (lane & 2) != 0;
    float n2[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float keep = b0 ? v[2 * q + 1] : v[2 * q];
        const float send = b0 ? v[2 * q] : v[2 * q + 1];
        n2[q] = fadd(keep, __shfl_xor_sync(FULL, send, 1));
    }
    const float keep = b1 ? n2[1] : n2[0];
    const float send = b1 ? n2[0] : n2[1];
    float r = fadd(keep, __shfl_xor_sync(FULL, send, 2));
    r = fadd(r, __shfl_xor_sync(FULL, r, 4));
    return r;
}

// phase profile that compiles to nothing in the production instantiation (12 x 64-bit accumulators cost 24 registers)
template <bool ON>
struct ProfT;
template <>
struct ProfT<true> : Prof {
    static constexpr bool on = true;
    __device__ __forceinline__ ProfT(long long *s) : Prof(s) {}
};
template <>
struct ProfT<false> {
    static constexpr bool on = false;
    long long acc[1];
    __device__ __forceinline__ ProfT(long long *) {}
    __device__ __forceinline__ void start() {}
    __device__ __forceinline__ void mark(int) {}
    __device__ __forceinline__ void stamp(int) {}
};

// shared-memory layout of a v2 layer CTA, in floats from g_smem
template <class SH>
struct V2L {
    using Cur = typename SH::Cur;
    using Lc = typename SH::Lc;
    using Gc = typename SH::Gc;
    using Skip = typename SH::Skip;
    static constexpr int BIAS = 2 * SH::Dm + SH::R + SH::Sm;                 // bfg | bd | bs
    static constexpr int WOLD = Cur::NPASS * Cur::CH * WN_NT, WLC = Lc::NPASS * Lc::CH * WN_NT, WGC = Gc::NPASS * Gc::CH * WN_NT;
    static constexpr int OFF_BFG = 0, OFF_BD = 2 * SH::Dm, OFF_BS = OFF_BD + SH::R;
    static constexpr int OFF_WOLD = BIAS, OFF_WLC = OFF_WOLD + WOLD, OFF_WGC = OFF_WLC + WLC;
    static constexpr int OFF_XS = OFF_WGC + WGC;                              // chain: layer input of the row in flight
    static constexpr int XSH = SH::R / 2 + 4;                                 // K-half stride of xs: the two halves fall in different banks
    static constexpr int OFF_ZS = OFF_XS + 2 * XSH;                           // chain: own gated slice for the dense (2 buffers)
    static constexpr int OFF_XSOLD = OFF_ZS + 2 * SH::Dm;                     // helper: dilated tap, padded for Cur
    static constexpr int OFF_LCS = OFF_XSOLD + Cur::TPC * Cur::XS;            // helper: lc row, padded for Lc
    static constexpr int OFF_GVEC = OFF_LCS + Lc::TPC * Lc::XS;               // prologue: speaker embedding
    static constexpr int OFF_WC = (OFF_GVEC + Gc::TPC * Gc::XS + 3) & ~3;     // layer 0: causal kernel [ifw][R]
    static constexpr int OFF_PART = OFF_WC + SH::IFW * SH::R;                 // layer 0: [Mt][32] conv2 partials of the row being drawn
    static constexpr int OFF_BAR = OFF_PART + SH::Mt * 32;                    // mbarriers (8 B each): 5 x 32 rows + image
    static constexpr int OFF_ROWS = OFF_BAR + 2 * (6 * WN_MAX_BATCH + 2);
    // per row
    static constexpr int ZF = Skip::TPC * Skip::XS;                           // full gated vector, padded for Skip
    static constexpr int R_INX = 0;                                            // [M][R] partial inputs (DSMEM inbox)
    static constexpr int R_HPRE = R_INX;                                       // layer 0 (no inbox): [4][R] causal partial sums
    static constexpr int R_Z = SH::M * SH::R;                                  // [ZF]
    static constexpr int R_ACC = R_Z + ZF;                                     // [Sm] running skip sum of layer l-1
    static constexpr int R_CQRING = R_ACC;                                     // layer 0 (no acc input): causal queue ring [IFW]
    static constexpr int R_XIN = R_ACC + SH::IFW;                              // layer 0: network input x_in of the step
    static constexpr int R_SAMP = R_ACC + SH::IFW + 16;                        // layer 0: [nr_mix] Gumbel noise, logistic noise, forced input of the next draw
    static constexpr int R_XRAW = R_ACC + SH::Sm;                              // [R] combined layer input
    static constexpr int R_PRE = R_XRAW + SH::R;                               // [2 Dm] pre-activations of the next step
    static constexpr int R_BFGN = R_PRE + 2 * SH::Dm;                          // [2 Dm] bias + speaker contribution
    static constexpr int R_MEL = R_BFGN + 2 * SH::Dm;                          // [C] mel frame of the row (TMA bulk copy, one frame = hop steps)
    static constexpr int ROWF = R_MEL + ((SH::C + 3) & ~3);
    __host__ __device__ static constexpr int total_floats(int n_rows) { return OFF_ROWS + n_rows * ROWF; }
};

template <int N4>
__device__ __forceinline__ void v2_load_wreg(float4 (&wv)[N4], const float *packed_base, int slot)
{
    const float4 *w4 = reinterpret_cast<const float4 *>(packed_base) + slot;
#pragma unroll
    for (int i = 0; i < N4; ++i) wv[i] = __ldg(w4 + (size_t)i * WN_NT);
}

// ---- bounded waits whose state lives in registers ---------------------------------------------------------------
// (wn_kernel.cu's watchdog_check takes its state by reference into a noinline function, which pins it to local memory
//  and puts an LDL in front of every barrier.)  Returns the (initialised) start time, or -1 when the launch is aborting.
__device__ __noinline__ long long v2_watchdog(int32_t *status, long long t0)
{
    if (t0 == 0) t0 = clock64();
    if (ld_volatile_i32(status) != 0) return -1;
    if (clock64() - t0 > WATCHDOG_CYCLES) {
        if (atomicCAS(status, 0, 1) == 0) { status[1] = 1000 + (int)blockIdx.x; status[2] = (int)threadIdx.x; }
        return -1;
    }
    return t0;
}
struct V2Ab {
    int32_t *status;
    int dead;          // this thread has seen the abort: every later wait falls through (the data is garbage from then on)
};
__device__ __forceinline__ void v2_mbar_wait(uint64_t *bar, unsigned parity, V2Ab &ab)
{
    if (mbar_try(bar, parity) || ab.dead) return;
    unsigned spins = 0;
    long long t0 = 0;
    while (!mbar_try(bar, parity)) {
        if (((++spins) & 0xffu) == 0) {
            t0 = v2_watchdog(ab.status, t0);
            if (t0 < 0) { ab.dead = 1; break; }
        }
    }
}
__device__ __forceinline__ float v2_ll_wait(const MBox &mb, const u64 *logical, unsigned seq, V2Ab &ab)
{
    const u64 *p = mb.rd(logical);
    u64 v = ld_relaxed_u64(p);
    unsigned spins = 0;
    long long t0 = 0;
    while ((unsigned)(v >> 32) != seq && !ab.dead) {
        if (((++spins) & 0x3ffu) == 0) {
            t0 = v2_watchdog(ab.status, t0);
            if (t0 < 0) { ab.dead = 1; break; }
        }
        v = ld_relaxed_u64(p);
    }
    return __uint_as_float((unsigned)v);
}
// n (<= 4) words p[i*stride], loads issued together
__device__ __forceinline__ void v2_ll_wait_n(const MBox &mb, const u64 *logical, size_t stride, int n, unsigned seq, V2Ab &ab, float *out)
{
    const u64 *pp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pp[i] = (i < n) ? mb.rd(logical + i * stride) : nullptr;
    u64 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (i < n) ? ld_relaxed_u64(pp[i]) : ((u64)seq << 32);
    unsigned spins = 0;
    long long t0 = 0;
    while (!ab.dead) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 4; ++i) ok = ok && ((unsigned)(v[i] >> 32) == seq);
        if (ok) break;
        if (((++spins) & 0x3ffu) == 0) {
            t0 = v2_watchdog(ab.status, t0);
            if (t0 < 0) { ab.dead = 1; break; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < n && (unsigned)(v[i] >> 32) != seq) v[i] = ld_relaxed_u64(pp[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = __uint_as_float((unsigned)v[i]);
}
// volatile loads: issued where they are written (prefetches for the NEXT item must leave before this item's waits)
__device__ __forceinline__ float v2_ld_cg_f32(const float *p)
{
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// one mel frame (C floats) global -> shared by a TMA bulk copy, completion on `bar` (one thread)
__device__ __forceinline__ void v2_mel_load(float *dst, const float *src, uint32_t bytes, uint64_t *bar)
{
    v2_expect_tx(bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// plain named barrier over the 256-thread helper group
__device__ __forceinline__ void v2_group_sync(int id)
{
    asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory");
}
// the 128-thread chain group: barrier 1
__device__ __forceinline__ void v2_chain_sync()
{
    asm volatile("bar.sync 1, 128;" ::: "memory");
}
__device__ __forceinline__ int v2_chain_sync_or(int flag)
{
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "barrier.red.or.pred q, 1, 128, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(flag)
        : "memory");
    return r;
}

// tanh / sigmoid through MUFU.EX2 + MUFU.RCP (~45 cycles instead of ~180 for the pinned exp32 + IEEE divide).
// Only for the scalar-input (mixture-of-logistics) path, whose tolerance is 1e-4 on the logits (north_star); the
// mu-law path keeps the pinned arithmetic because its integer samples must be bit-exact.  |error| <= ~3e-7.
__device__ __forceinline__ float act_fg_fast(float x, bool is_gate)
{
    const float arg = fmul(is_gate ? -x : fadd(x, x), 1.4426950408889634f);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fadd(e, 1.0f)));
    return is_gate ? r : ffma(-2.0f, r, 1.0f);
}

// Mixture-of-logistics draw from a warp's logits (lane o < O holds logit o): wavenet/mixture.py:84-114.  Returns the
// sample in every lane; `writer` stores it.
template <class SH>
__device__ __forceinline__ float v2_draw_warp(const WnParams &p, int b, int step, int lane, bool writer, float c2, float gum, float logistic)
{
    constexpr int nr = SH::O / 3;
    float g = (lane < nr) ? fsub(c2, gum) : __int_as_float(0xff800000);
    int k = lane;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float og = __shfl_xor_sync(FULL, g, off);
        const int ok = __shfl_xor_sync(FULL, k, off);
        if (og > g || (og == g && ok < k)) { g = og; k = ok; }
    }
    const float mean = __shfl_sync(FULL, c2, nr + k);
    float ls = __shfl_sync(FULL, c2, 2 * nr + k);
    const float lsmin = -32.23619130191664f;
    if (!(ls > lsmin)) ls = lsmin;
    float x = fadd(mean, fmul(wn::exp32(ls), logistic));
    x = fmaxf(x, -1.0f);
    x = fminf(x, 1.0f);
    if (writer && lane == 0) p.out_samples[(size_t)b * p.T + step] = x;
    return x;
}

// One warp of a layer-0 CTA: waits for the Mt partial conv2 outputs of (b, step) from the tail CTAs, adds them left to
// right onto the bias, writes the logits, draws from the mixture of logistics (wavenet/mixture.py:84-114) and returns
// the sample in every lane.  Lane o < O owns output o; gum / logistic are precomputed from the step's uniforms.
template <class SH, class PF>
__device__ __forceinline__ float v2_sample_warp(const WnParams &p, const MBox &mb, int b, int step, int lane, bool writer, V2Ab &ab,
                                                float b2v, float gum, float logistic, PF &pf)
{
    constexpr int O = SH::O, Mt = SH::Mt, nr = SH::O / 3;
    static_assert(O <= 32 && Mt == 16, "sample warp: one lane per output, 16 tail partials");
    const unsigned seq = (unsigned)step + 1u;
    float c2 = b2v;
    unsigned spins = 0;
    if (lane < O) {
        // 16 words per lane, all in flight; consumed in order (the partial sums are added left to right).  A missing word
        // re-issues the loads of every word not yet consumed, so late tail CTAs cost one poll round, not one each.
        // The lane's 16 words (stride O) span at most three 2 KB grains of the logical mailbox space: resolve those three
        // grain bases once instead of a table lookup per word and per poll round.
        const size_t w0 = (reinterpret_cast<size_t>(p.mb_c2) >> 3) + ((size_t)b * Mt) * O + lane;
        const size_t g0 = w0 >> 8;
        static_assert((Mt - 1) * O < 512, "sample warp: the partial words of one output fit three grains");
        const unsigned long long *tab = reinterpret_cast<const unsigned long long *>(mb.mine);
        const u64 *gb0 = reinterpret_cast<const u64 *>(__ldg(tab + g0));
        const u64 *gb1 = reinterpret_cast<const u64 *>(__ldg(tab + g0 + 1));
        const u64 *gb2 = reinterpret_cast<const u64 *>(__ldg(tab + g0 + 2));
        auto word = [&](int i) -> const u64 * {
            const size_t w = w0 + (size_t)i * O;
            const size_t g = (w >> 8) - g0;
            const u64 *base = (g == 0) ? gb0 : ((g == 1) ? gb1 : gb2);
            return base + (w & 255);
        };
        u64 wv[Mt];
#pragma unroll
        for (int i = 0; i < Mt; ++i) wv[i] = ld_relaxed_u64(word(i));
        long long t0 = 0;
#pragma unroll
        for (int i = 0; i < Mt; ++i) {
            while ((unsigned)(wv[i] >> 32) != seq && !ab.dead) {
#pragma unroll
                for (int j = i; j < Mt; ++j) wv[j] = ld_relaxed_u64(word(j));
                if (((++spins) & 0x3ffu) == 0) {
                    t0 = v2_watchdog(ab.status, t0);
                    if (t0 < 0) { ab.dead = 1; break; }
                }
            }
            c2 = fadd(c2, __uint_as_float((unsigned)wv[i]));
        }
        if (writer && p.out_logits) p.out_logits[((size_t)b * p.T + step) * O + lane] = c2;
    }
    __syncwarp();
    return v2_draw_warp<SH>(p, b, step, lane, writer, c2, gum, logistic);
}

// SR ("shared ring", launches with >= 16 rows): ONE dilation ring per layer instead of one private copy per sibling CTA.  The
// four copies (1.59 MB x 4 per row) overflow the L2 from about 10 rows on (profiles/r02_ncu_range_16x12000.csv: 3.9 GB of DRAM
// traffic per launch at 16 rows) and put HBM round trips into the helper groups' taps.  Every sibling holds the full layer input, so
// sibling m pushes only its quarter of x; the ring holds 8-byte {value, step tag} words (the LL mailbox protocol) and a reader
// that finds an older tag re-reads until the sibling's push has landed -- no ordering between the siblings is assumed.  The ring is
// zeroed per launch (tag 0 is never expected).  SR = false is the unchanged round-2 code: the 8-row headline path keeps its SASS.
template <class SH, bool FAST, bool PROF, int CS, bool SR>
__device__ void layer_role_v2(const WnParams &p, const int l, const int m, const int l_local)
{
    using Cur = typename SH::Cur;        // packed for 256 slots: col = slot / 4, chunk = slot % 4, 8 float4
    using Lc = typename SH::Lc;
    using Gc = typename SH::Gc;
    using Dense = typename SH::Dense;    // col = slot / 2, chunk = slot % 2, 4 float4
    using Skip = typename SH::Skip;      // col = slot / 2, chunk = slot % 2, 16 float4
    using LY = V2L<SH>;
    static_assert(Cur::TPC == 4 && Cur::N4 == 8 && Cur::U == 8 && Cur::NPASS == 1, "v2 mapping assumes the cfg-2 fg shape");
    static_assert(Dense::TPC == 2 && Dense::N4 == 4 && Dense::U == 4 && Dense::NPASS == 1, "v2 mapping assumes the cfg-2 dense shape");
    static_assert(Skip::TPC == 2 && Skip::NPASS == 1 && Lc::TPC == 4 && Gc::TPC == 4, "v2 mapping assumes the cfg-2 shapes");
    constexpr int R = SH::R, M = SH::M, Dm = SH::Dm, Sm = SH::Sm, D = SH::D, ncol2 = 2 * SH::Dm;
    static_assert(R == 128 && D == 128 && Sm == 128 && M == 4 && Dm == 32, "v2 mapping: 128-wide vectors, 4 CTAs per layer");
    static_assert(SH::SCALAR, "v2 layer kernel is instantiated for scalar-input models");

    float *smem = g_smem;
    const int tid = threadIdx.x;
    const int N = p.N, L = p.L;
    const int cta = l * M + m;
    const float *gimg = p.layer_img + (size_t)cta * p.layer_img_floats;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + LY::OFF_BAR);
    uint64_t *xbar = bars, *zbar = bars + WN_MAX_BATCH, *abar = bars + 2 * WN_MAX_BATCH, *fullb = bars + 3 * WN_MAX_BATCH,
             *prdy = bars + 4 * WN_MAX_BATCH, *melbar = bars + 5 * WN_MAX_BATCH, *ldbar = bars + 6 * WN_MAX_BATCH;
    float *rows = smem + LY::OFF_ROWS;

    constexpr int LPC = CS / 4;                               // consecutive layers per cluster
    const unsigned lbase = (unsigned)l_local * 4u;            // cluster rank of this layer's CTA 0
    const unsigned nbase = lbase + 4u;                        // ... and of the next layer's, when it is a cluster mate
    const bool in_dsmem = l_local > 0;                        // fed by a cluster mate through distributed shared memory
    const bool out_dsmem = l_local + 1 < LPC && l + 1 < L && l + 1 < p.layer_end;
    const bool has_next_layer = l + 1 < L;

    // ---- barriers, resident image (TMA bulk copies), scratch --------------------------------------------------
    if (tid == 0) {
        for (int b = 0; b < WN_MAX_BATCH; ++b) {
            mbar_init(&xbar[b], 1); mbar_init(&zbar[b], 1); mbar_init(&abar[b], 1); mbar_init(&fullb[b], 1); mbar_init(&prdy[b], 1);
            mbar_init(&melbar[b], 1);
        }
        mbar_init(ldbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes1 = (uint32_t)LY::BIAS * 4u, bytes2 = (uint32_t)(LY::WOLD + LY::WLC + LY::WGC) * 4u;
        v2_expect_tx(ldbar, bytes1 + bytes2);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
                     "l"(gimg + p.off_bfg), "r"(bytes1), "r"(smem_u32(ldbar))
                     : "memory");
        const uint32_t CHK = 30720u;
        for (uint32_t o = 0; o < bytes2; o += CHK) {
            const uint32_t sz = (bytes2 - o < CHK) ? (bytes2 - o) : CHK;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32((const char *)(smem + LY::OFF_WOLD) + o)),
                         "l"((const char *)(gimg + p.old.off) + o), "r"(sz), "r"(smem_u32(ldbar))
                         : "memory");
        }
    }
    for (int i = LY::OFF_XS + tid; i < LY::OFF_WC; i += V2_NT) smem[i] = 0.0f;
    for (int i = tid; i < N * LY::ROWF; i += V2_NT) rows[i] = 0.0f;
    if (l == 0) {
        // causal kernel (wavenet/model.py:41-46) as [tap][channel]; packed in the sampler image for 256 slots (Causal: TPC 2, 4 float4)
        using Causal = typename SH::Causal;
        static_assert(Causal::TPC == 2 && Causal::N4 == 4 && Causal::U == 4 && SH::IFW == 32, "v2 causal layer assumes the cfg-2 shape");
        const float *wcp = p.samp_img + p.causal.off;
        for (int i = tid; i < SH::IFW * R; i += V2_NT) {
            const int k = i / R, r = i % R;
            smem[LY::OFF_WC + i] = __ldg(wcp + (((size_t)((k % 16) / 4) * WN_NT + r * 2 + k / 16) * 4 + (k % 4)));
        }
    }
    while (!mbar_try(ldbar, 0)) {}
    __syncthreads();
    if (tid == 0) {
        for (int b = 0; b < N; ++b) {
            if (in_dsmem) v2_expect_tx(&xbar[b], (uint32_t)(M * R * 4));
            v2_expect_tx(&zbar[b], (uint32_t)((M - 1) * Dm * 4));
            if (in_dsmem) v2_expect_tx(&abar[b], (uint32_t)(Sm * 4));
        }
    }
    // every CTA of the cluster has initialised and armed its barriers before anyone stores into a peer
    v2_cluster_sync();

    const float *bfg = smem + LY::OFF_BFG, *bd = smem + LY::OFF_BD, *bs = smem + LY::OFF_BS;
    float *xs = smem + LY::OFF_XS, *zs = smem + LY::OFF_ZS, *xs_old = smem + LY::OFF_XSOLD, *lcs = smem + LY::OFF_LCS,
          *gvec = smem + LY::OFF_GVEC;
    const float *w_old_s = smem + LY::OFF_WOLD, *w_lc_s = smem + LY::OFF_WLC, *w_gc_s = smem + LY::OFF_WGC;

    const int d = p.dil[l];
    float *ring_cta = p.ring + p.ring_off[l] + (size_t)m * N * d * R;
    u64 *ring_sh = reinterpret_cast<u64 *>(p.ring + p.ring_off[l]);            // SR: (N, d, R) tagged words = half of the layer's region
    V2Ab ab{p.status, 0};
    const MBox mb = make_mbox(p);
    const size_t rowx = (size_t)L * M * R, rowa = (size_t)L * M * Sm;

    if (tid < V2_HALF) {
        // =========================== HELPER group (warps 0-7) ===================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(V2_REGS_HELPER));
        const int ht = tid;
        const int c4 = ht & 3, c2 = ht & 1;
        const int grp4 = ht >> 2, col2 = ht >> 1;
        const bool lead4 = c4 == 0, lead2 = c2 == 0;
        const float *xc_old = xs_old + c4 * Cur::XS;
        const float *xc_lc = lcs + c4 * Lc::XS;
        const float *xc_gc = gvec + c4 * Gc::XS;
        const float *w_old = w_old_s + ht * 4, *w_lc = w_lc_s + ht * 4, *w_gc = w_gc_s + ht * 4;
        const int xp_x = (ht < R) ? Cur::xpad(ht) : 0;
        const int xp_lc = (SH::HAS_LC && ht < SH::C) ? Lc::xpad(ht) : 0;
        float4 wsk[Skip::N4];                                   // skip 1x1 slice: 64 registers, resident
        v2_load_wreg<Skip::N4>(wsk, gimg + p.skip.off, ht);
        const float bsv = bs[col2];
        ProfT<PROF> hp((p.prof && ht == 0) ? p.prof + (size_t)cta * 16 : nullptr);

        // ---- prologue: speaker contribution folded into the biases, pre-activations for t = 0 ------------------
        for (int b = 0; b < N; ++b) {
            float *rb = rows + (size_t)b * LY::ROWF;
            float *bfgN = rb + LY::R_BFGN, *pre_b = rb + LY::R_PRE;
            if (SH::HAS_GC) {
                if (ht < SH::G) gvec[Gc::xpad(ht)] = __ldg(p.gc_table + (size_t)p.gc_id[b] * SH::G + ht);
                v2_group_sync(2);
                matvec_s<Gc>(w_gc, xc_gc, grp4, lead4, ncol2, [&](int col, float dot) { bfgN[col] = fadd(bfg[col], dot); });
            } else {
                if (ht < ncol2) bfgN[ht] = bfg[ht];
            }
            v2_group_sync(2);
            // dilated tap = zeros, lc = zeros at t = 0 (xs_old / lcs are zero-initialised)
            matvec_s<Cur>(w_old, xc_old, grp4, lead4, ncol2, [&](int col, float dot) { pre_b[col] = fadd(bfgN[col], dot); });
            if (SH::HAS_LC) {
                v2_group_sync(2);
                matvec_s<Lc>(w_lc, xc_lc, grp4, lead4, ncol2, [&](int col, float dot) { pre_b[col] = fadd(pre_b[col], dot); });
            }
            v2_group_sync(2);
        }
        if (l == 0 && ht < N && p.T_row[ht] > 0) rows[(size_t)ht * LY::ROWF + LY::R_SAMP + SH::O / 3 + 1] = __ldg(p.forced + (size_t)ht * p.n_forced);
        // folded create_upsample (model.py:102-111): frame 0 of every row, staged by a TMA bulk copy (phase 0 of melbar[b])
        const bool fold_lc = SH::HAS_LC && p.mel != nullptr;
        if (fold_lc && ht == 0)
            for (int b = 0; b < N; ++b)
                if (p.T_row[b] > 0) v2_mel_load(rows + (size_t)b * LY::ROWF + LY::R_MEL, p.mel + (size_t)b * p.t_mel * SH::C, SH::C * 4, &melbar[b]);
        v2_group_sync(2);
        if (ht == 0)
            for (int b = 0; b < N; ++b) mbar_arrive(&prdy[b]);           // phase 0: pre for t = 0 is ready

        // acc hand-over to layer l+1, CTA m: DSMEM inside the cluster, LL mailbox otherwise (and to the tail)
        const u64 *mba_in = p.mb_acc + ((size_t)(l > 0 ? l - 1 : 0) * M + m) * Sm + col2;
        const u64 *mba_out = p.mb_acc + ((size_t)l * M + m) * Sm + col2;
        const uint32_t r_acc = v2_mapa(smem_u32(rows + LY::R_ACC + col2), (out_dsmem ? nbase : lbase) + (unsigned)m);
        const uint32_t r_abar = v2_mapa(smem_u32(&abar[0]), (out_dsmem ? nbase : lbase) + (unsigned)m);
        const float4 *xc_skip = reinterpret_cast<const float4 *>(rows + LY::R_Z + c2 * Skip::XS);

        // Global loads of an item (dilated tap from the ring, materialised lc row, layer 0: the draw's uniforms / forced input).
        // They are issued one item AHEAD (while the current item waits and computes): in a train of rows the helper is the
        // busiest group of a CTA, and 600-1000 cycles of exposed L2 / DRAM latency per item would be added to its service time.
        struct GLoad { float oldv, lcraw, su; u64 oldw; };
        auto gload = [&](int t, int b) -> GLoad {
            GLoad g{0.0f, 0.0f, 0.0f, 0ull};
            const bool has_next = (t + 1 < p.T_row[b]);
            if (has_next && d >= 2 && t + 1 >= d && ht < R) {
                if constexpr (SR) g.oldw = ld_relaxed_u64(ring_sh + ((size_t)b * d + ((t + 1) % d)) * R + ht);
                else g.oldv = v2_ld_cg_f32(ring_cta + ((size_t)b * d + ((t + 1) % d)) * R + ht);       // x_l(t+1-d); the queue starts at zero (model.py:64)
            }
            if (SH::HAS_LC && !fold_lc && has_next && ht < SH::C && p.lc_up != nullptr) {
                const long idx = (long)t - p.lc_shift;
                if (idx >= 0 && idx < p.t_lc) g.lcraw = ld_nc_f32(p.lc_up + ((size_t)b * p.t_lc + idx) * SH::C + ht);
            }
            if (l == 0 && has_next && ht <= SH::O / 3 + 1) {
                constexpr int nr = SH::O / 3;
                if (ht <= nr) g.su = ld_nc_f32(p.noise + ((size_t)b * p.T + t) * (nr + 1) + ht);
                else if (t + 1 < p.n_forced) g.su = ld_nc_f32(p.forced + (size_t)b * p.n_forced + t + 1);
            }
            return g;
        };
        GLoad pref{0.0f, 0.0f, 0.0f, 0ull};
        int pref_t = -1, pref_b = -1;

        for (int t = 0; t < p.T; ++t) {
            // abort is agreed on by the whole group, at a step boundary (every 16th, to keep it off the per-row path)
            if ((t & 15) == 0 && v2_group_sync_or(2, ab.dead)) break;
            const unsigned seq = (unsigned)t + 1u;
            const unsigned par = (unsigned)t & 1u;
            for (int b = 0; b < N; ++b) {
                if (t >= p.T_row[b]) continue;
                hp.start();
                float *rb = rows + (size_t)b * LY::ROWF;
                const bool has_next = (t + 1 < p.T_row[b]);
                MDst da{nullptr, nullptr};
                if (!out_dsmem) da = mb_dst(mb, mba_out + b * rowa);
                const GLoad cur = (pref_t == t && pref_b == b) ? pref : gload(t, b);
                {
                    // the item after this one; its loads leave now unless it is the same row (its ring slot may be this item's push)
                    int tn = t, bn = b;
                    bool hn = false;
                    for (int k = 0; k <= N; ++k) {
                        if (++bn >= N) { bn = 0; ++tn; }
                        if (tn >= p.T) break;
                        if (tn < p.T_row[bn]) { hn = true; break; }
                    }
                    pref_t = -1;
                    if (hn && bn != b) { pref = gload(tn, bn); pref_t = tn; pref_b = bn; }
                }
                float oldv = cur.oldv, lcv = cur.lcraw;
                if constexpr (SR) {
                    if (has_next && d >= 2 && t + 1 >= d && ht < R) {
                        // x_l(t+1-d) was pushed with tag (t+1-d) + 1 by the sibling that owns channel ht
                        const unsigned want = (unsigned)(t + 2 - d);
                        const u64 *pw = ring_sh + ((size_t)b * d + ((t + 1) % d)) * R + ht;
                        u64 wv = cur.oldw;
                        unsigned spins = 0;
                        long long t0 = 0;
                        while ((unsigned)(wv >> 32) != want && !ab.dead) {
                            if (((++spins) & 0x3ffu) == 0) {
                                t0 = v2_watchdog(ab.status, t0);
                                if (t0 < 0) { ab.dead = 1; break; }
                            }
                            wv = ld_relaxed_u64(pw);
                        }
                        oldv = __uint_as_float((unsigned)wv);
                    }
                }
                bool mel_last = false;
                int mel_next = 0;
                if (SH::HAS_LC && has_next && ht < SH::C) {
                    long idx = (long)t - p.lc_shift;
                    if (fold_lc) {
                        // row idx of the upsampled condition = three conv2d_transpose stages of mel frame idx / hop, evaluated
                        // exactly as wn_upsample_stage_kernel does (fmul, then fma with the left neighbour), positions < 0 are 0
                        if (idx >= 0 && idx < (long)p.t_mel * p.hop) {
                            const int fr = (int)(idx / p.hop);
                            int rem = (int)(idx % p.hop);
                            const int a2 = rem % p.up_f[2]; rem /= p.up_f[2];
                            const int a1 = rem % p.up_f[1];
                            const int a0 = rem / p.up_f[1];
                            if (idx % p.hop == 0) v2_mbar_wait(&melbar[b], (unsigned)fr & 1u, ab);      // frame fr has landed
                            mel_last = (idx % p.hop == p.hop - 1) && fr + 1 < p.t_mel;
                            mel_next = fr + 1;
                            const float *mf = rb + LY::R_MEL;
                            float v[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] = (ht - j >= 0) ? mf[ht - j] : 0.0f;
                            const int aa[3] = {a0, a1, a2};
#pragma unroll
                            for (int sgi = 0; sgi < 3; ++sgi) {
                                const float k0 = __ldg(p.upk + p.up_off[sgi] + aa[sgi] * 2), k1 = __ldg(p.upk + p.up_off[sgi] + aa[sgi] * 2 + 1);
#pragma unroll
                                for (int j = 0; j < 3 - sgi; ++j) v[j] = (ht - j >= 0) ? ffma(v[j + 1], k1, fmul(v[j], k0)) : 0.0f;
                            }
                            lcv = v[0];
                        }
                    }
                }
                // layer 0: noise of the draw of step t and the forced input of step t+1 (consumed by the chain's item (b, t+1))
                float sprep = 0.0f;
                if (l == 0 && has_next && ht <= SH::O / 3 + 1) {
                    constexpr int nr = SH::O / 3;
                    sprep = cur.su;                                      // Gumbel / logistic noise (wn_noise_prep_kernel) or the forced input
                }
                pin(oldv); pin(lcv); pin(sprep);
                hp.mark(6);
                v2_mbar_wait(&fullb[b], par, ab);                       // x and the own z slice of (b, t) are in rows[b]
                if (l == 0 && has_next && ht <= SH::O / 3 + 1) rb[LY::R_SAMP + ht] = sprep;
                hp.mark(7);
                if (ht < R) {
                    const float xme = rb[LY::R_XRAW + ht];
                    if constexpr (SR) {
                        if (d >= 2 && ht / (R / M) == m) ll_store(ring_sh + ((size_t)b * d + (t % d)) * R + ht, xme, seq);
                    } else {
                        if (d >= 2) __stcg(ring_cta + ((size_t)b * d + (t % d)) * R + ht, xme);
                    }
                    if (d == 1) oldv = xme;
                    xs_old[xp_x] = oldv;
                }
                if (SH::HAS_LC && ht < SH::C) lcs[xp_lc] = lcv;
                if (l == 0 && ht == 0) rb[LY::R_CQRING + (t & (SH::IFW - 1))] = rb[LY::R_XIN];     // causal queue push (model.py:122)
                v2_mbar_wait(&zbar[b], par, ab);                        // the three siblings' slices have landed
                if (ht == 0) v2_expect_tx(&zbar[b], (uint32_t)((M - 1) * Dm * 4));
                v2_group_sync(2);
                // every thread has consumed the row's mel frame: fetch the next one, it is first read a full step from now
                if (SH::HAS_LC && ht == 0 && mel_last)
                    v2_mel_load(rb + LY::R_MEL, p.mel + ((size_t)b * p.t_mel + mel_next) * SH::C, SH::C * 4, &melbar[b]);
                hp.mark(8);
                // skip 1x1 column col2 over K half c2, from registers, + running skip sum
                {
                    const float4 *zc = xc_skip + (size_t)b * (LY::ROWF / 4);
                    constexpr int PER = Skip::N4 / Skip::U;
                    float acc[Skip::U];
#pragma unroll
                    for (int u = 0; u < Skip::U; ++u) {
                        float a = 0.0f;
#pragma unroll
                        for (int i = 0; i < PER; ++i) {
                            const float4 ww = wsk[u * PER + i], xx = zc[u * PER + i];
                            a = ffma(ww.x, xx.x, a);
                            a = ffma(ww.y, xx.y, a);
                            a = ffma(ww.z, xx.z, a);
                            a = ffma(ww.w, xx.w, a);
                        }
                        acc[u] = a;
                    }
#pragma unroll
                    for (int off = 1; off < Skip::U; off <<= 1)
#pragma unroll
                        for (int c = 0; c < Skip::U; c += 2 * off) acc[c] = fadd(acc[c], acc[c + off]);
                    const float dsk = butterfly<2>(acc[0]);
                    if (lead2) {
                        float v = fadd(bsv, dsk);
                        if (l > 0) {
                            float a;
                            if (in_dsmem) {
                                v2_mbar_wait(&abar[b], par, ab);
                                a = rb[LY::R_ACC + col2];
                            } else {
                                a = v2_ll_wait(mb, mba_in + b * rowa, seq, ab);
                            }
                            v = fadd(a, v);
                        }
                        if (out_dsmem) v2_st_async(r_acc + (uint32_t)b * (LY::ROWF * 4), v, r_abar + (uint32_t)b * 8u);
                        else ll_post(da, v, seq);
                    }
                }
                hp.mark(9);
                // layer 0: the 31 known taps of the next step's causal conv (model.py:131), as the partial sums the chain
                // thread combines with the new sample: canonical chunks of 4 taps, a[c] += a[c^1], a[c^2], a[c^4]
                if (l == 0 && has_next && ht < R) {
                    const float *wc = smem + LY::OFF_WC + ht;
                    const float *cq = rb + LY::R_CQRING;
                    float a[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float v = 0.0f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int k = 4 * c + i;                     // tap k holds x_in(t - 30 + k); tap 31 is the sample to come
                            if (k < SH::IFW - 1) v = ffma(wc[k * R], cq[(t + 2 + k) & (SH::IFW - 1)], v);
                        }
                        a[c] = v;
                    }
                    float *hp4 = rb + LY::R_HPRE + ht;
                    hp4[0] = fadd(fadd(a[0], a[1]), fadd(a[2], a[3]));
                    hp4[R] = fadd(a[4], a[5]);
                    hp4[2 * R] = a[6];
                    hp4[3 * R] = a[7];
                }
                // pre-activations of the next step: (bias(+gc) + W_old . x_l(t+1-d)) + W_lc . lc(t)
                if (has_next) {
                    float *pre_b = rb + LY::R_PRE;
                    const float *bias_b = rb + LY::R_BFGN;
                    float pv = butterfly<4>(dot_regs<Cur::N4, Cur::U>(reinterpret_cast<const float4 *>(w_old), reinterpret_cast<const float4 *>(xc_old)));
                    pv = fadd(bias_b[grp4], pv);
                    if (SH::HAS_LC)
                        pv = fadd(pv, butterfly<4>(dot_regs<Lc::N4, Lc::U>(reinterpret_cast<const float4 *>(w_lc), reinterpret_cast<const float4 *>(xc_lc))));
                    if (lead4) pre_b[grp4] = pv;
                }
                v2_group_sync(2);
                // all lead2 threads are past their abar wait: re-arm it for the next step
                if (ht == 0) {
                    if (in_dsmem) v2_expect_tx(&abar[b], (uint32_t)(Sm * 4));
                    mbar_arrive(&prdy[b]);                               // pre[b] for t+1 written, rows[b] free
                }
                hp.mark(10);
            }
        }
        if (PROF && p.prof && ht == 0)
            for (int i = 7; i < 11; ++i) p.prof[(size_t)cta * 16 + i] = hp.acc[PROF ? i : 0];
    } else {
        // =========================== CHAIN group (warps 8-11) ===================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(V2_REGS_CHAIN));
        // "Fat" threads, one warp per SM sub-partition: thread = (filter/gate column c, K half) holds 64 weights of the
        // current tap and evaluates canonical chunks khalf*16 .. khalf*16+15 (in-thread tree, ONE shuffle level);
        // for the dense 1x1 thread = output r over the CTA's whole 32-channel slice (8 canonical chunks, no shuffle).
        // Two instantiations of the same body: the layer-0 one carries the sampler (16 LL words in flight per lane), whose
        // register demand would otherwise spill the weight registers of every layer's loop.
        auto chain_group = [&](auto l0_tag) {
        constexpr bool L0 = decltype(l0_tag)::value;
        const int ct = tid - V2_HALF;
        const int c = ct >> 1, khalf = ct & 1;                       // column c = 2*j + gate
        const bool gate = (c & 1) != 0;
        const int zj = c >> 1;                                       // gated channel inside the CTA
        float4 wfg[16];
        {
            // packed slot (wn_params.h, Cur: TPC 4, 8 float4): col*4 + k/32, float4 index (k%32)/4
            const float4 *pk = reinterpret_cast<const float4 *>(gimg + p.cur.off);
#pragma unroll
            for (int i = 0; i < 16; ++i) wfg[i] = __ldg(pk + (size_t)(i & 7) * WN_NT + c * 4 + khalf * 2 + (i >> 3));
        }
        float4 wdn[8];
        {
            // Dense: TPC 2, 4 float4: slot r*2 + k/16, float4 index (k%16)/4
            const float4 *pd = reinterpret_cast<const float4 *>(gimg + p.dense.off);
#pragma unroll
            for (int i = 0; i < 8; ++i) wdn[i] = __ldg(pd + (size_t)(i & 3) * WN_NT + ct * 2 + (i >> 2));
        }
        const float4 *xh4 = reinterpret_cast<const float4 *>(xs + khalf * LY::XSH);
        const int xp = (ct >> 6) * LY::XSH + (ct & 63);              // where combined input ct goes in xs
        // z hand-over: of the four threads that hold z_j, thread q = 0 keeps it, q = 1..3 ship it to sibling (m + q) & 3
        const unsigned zq = (unsigned)ct & 3u;
        const int zidx = Skip::xpad(m * Dm + zj);
        const uint32_t r_z = v2_mapa(smem_u32(rows + LY::R_Z + zidx), lbase + (((unsigned)m + zq) & 3u));
        const uint32_t r_zbar = v2_mapa(smem_u32(&zbar[0]), lbase + (((unsigned)m + zq) & 3u));
        const float bdv = bd[ct];
        uint32_t r_x[4], r_xb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r_x[k] = v2_mapa(smem_u32(rows + LY::R_INX + m * R + ct), (out_dsmem ? nbase : lbase) + (unsigned)k);
            r_xb[k] = v2_mapa(smem_u32(&xbar[0]), (out_dsmem ? nbase : lbase) + (unsigned)k);
        }
        const u64 *mbx_in = p.mb_x + ((size_t)l * M) * R + ct;
        const u64 *mbx_out = p.mb_x + ((size_t)(l + 1) * M + m) * R + ct;
        constexpr int nr_mix = SH::O / 3;
        const float w31r = smem[LY::OFF_WC + (SH::IFW - 1) * R + ct];                  // layer 0: newest causal tap
        const float b2v = (L0 && (ct & 31) < SH::O) ? __ldg(p.samp_img + p.off_b2 + (ct & 31)) : 0.0f;
        ProfT<PROF> pf((p.prof && ct == 0) ? p.prof + (size_t)cta * 16 : nullptr);
        unsigned item = 0;                                           // parity selects the zs buffer

        for (int t = 0; t < p.T; ++t) {
            if ((t & 15) == 0 && v2_chain_sync_or(ab.dead)) break;
            const unsigned seq = (unsigned)t + 1u;
            const unsigned par = (unsigned)t & 1u;
            for (int b = 0; b < N; ++b) {
                if (t >= p.T_row[b]) continue;
                pf.start();
                float *rb = rows + (size_t)b * LY::ROWF;
                float *zsb = zs + (item & 1u) * Dm;
                ++item;
                MDst dx{nullptr, nullptr};
                if (has_next_layer && !out_dsmem) dx = mb_dst(mb, mbx_out + b * rowx);
                v2_mbar_wait(&prdy[b], par, ab);                       // pre[b] of this step written, rows[b] free
                if (L0) pf.mark(0);                                    // layer 0: 'wait' = the helper's pre-activations, 'combine' = sampler + causal
                const float pre_v = rb[LY::R_PRE + c];
                // 1. layer input r = ct
                {
                    float v;
                    if (L0) {
                        // layer 0 hosts the sampler: warp 0 of the group draws sample t-1 from the tail's conv2 partials,
                        // the new network input closes the causal conv whose 31 older taps the helper has already summed
                        const float *hp4 = rb + LY::R_HPRE + ct;
                        const float h0 = hp4[0], h1 = hp4[R], h2 = hp4[2 * R], h3 = hp4[3 * R];
                        // Gumbel / logistic noise and the forced input were prepared by the helper one step ahead.
                        // Every warp of the group polls a quarter of the tail's 16 conv2 partials (4 words in flight per
                        // lane: one L2 round trip instead of four for 16 words from one warp), parks them in shared memory,
                        // and after the barrier every warp sums them in the pinned order and draws the sample redundantly:
                        // the new input reaches all 128 threads without a second barrier.
                        const int sl = ct & 31, cw = ct >> 5;
                        const float gum = (sl < nr_mix) ? rb[LY::R_SAMP + sl] : 0.0f;
                        const float logistic = rb[LY::R_SAMP + nr_mix];
                        float x_in = rb[LY::R_SAMP + nr_mix + 1];
                        float *part = smem + LY::OFF_PART;
                        pf.mark(8);
                        if (t > 0 && sl < SH::O) {
                            constexpr int PW = SH::Mt / 4;
                            float q[4];
                            v2_ll_wait_n(mb, p.mb_c2 + ((size_t)b * SH::Mt + cw * PW) * SH::O + sl, (size_t)SH::O, PW, (unsigned)t, ab, q);
#pragma unroll
                            for (int i = 0; i < PW; ++i) part[(cw * PW + i) * 32 + sl] = q[i];
                        }
                        pf.mark(7);
                        v2_chain_sync();
                        pf.mark(9);
                        if (t > 0) {
                            float c2 = b2v;
                            if (sl < SH::O) {
#pragma unroll
                                for (int i = 0; i < SH::Mt; ++i) c2 = fadd(c2, part[i * 32 + sl]);
                                if (m == 0 && cw == 0 && p.out_logits) p.out_logits[((size_t)b * p.T + (t - 1)) * SH::O + sl] = c2;
                            }
                            const float smp = v2_draw_warp<SH>(p, b, t - 1, sl, m == 0 && cw == 0, c2, gum, logistic);
                            if (t >= p.n_forced) x_in = smp;
                        }
                        if (ct == 0) rb[LY::R_XIN] = x_in;               // the helper pushes it into the causal queue
                        pf.mark(1);
                        pf.stamp(10);
                        v = fadd(h0, fadd(h1, fadd(h2, ffma(w31r, x_in, h3))));
                    } else if (in_dsmem) {
                        v2_mbar_wait(&xbar[b], par, ab);
                        pf.mark(0);
                        pf.stamp(10);
                        const float *in = rb + LY::R_INX + ct;
                        v = fadd(fadd(fadd(in[0], in[R]), in[2 * R]), in[3 * R]);
                    } else {
                        float q[4];
                        v2_ll_wait_n(mb, mbx_in + b * rowx, (size_t)R, M, seq, ab, q);
                        pf.mark(0);
                        pf.stamp(10);
                        v = fadd(fadd(fadd(q[0], q[1]), q[2]), q[3]);
                    }
                    xs[xp] = v;
                    rb[LY::R_XRAW + ct] = v;
                }
                pf.mark(1);
                v2_chain_sync();
                if (in_dsmem && ct == 0) v2_expect_tx(&xbar[b], (uint32_t)(M * R * 4));     // every reader is past its wait
                pf.mark(2);
                // 2. current-tap filter/gate column over one K half, from registers; gated activation
                {
                    // all 16 LDS.128 of the K half are issued before the first FMA, as volatile asm: left to itself ptxas sometimes
                    // splits them into two dependent batches of 8 (64 input + 96 weight registers are a tight fit), which costs
                    // ~120 cycles per layer (profiles/r02_rejected_variants.md: the "slow allocation" of two unrelated edits)
                    float4 xv[16];
                    {
                        const uint32_t xa = smem_u32(xh4);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(xv[i].x), "=f"(xv[i].y), "=f"(xv[i].z), "=f"(xv[i].w)
                                         : "r"(xa + (uint32_t)i * 16u));
                    }
                    float acc[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float a = 0.0f;
                        a = ffma(wfg[i].x, xv[i].x, a);
                        a = ffma(wfg[i].y, xv[i].y, a);
                        a = ffma(wfg[i].z, xv[i].z, a);
                        a = ffma(wfg[i].w, xv[i].w, a);
                        acc[i] = a;
                    }
#pragma unroll
                    for (int off = 1; off < 16; off <<= 1)
#pragma unroll
                        for (int k = 0; k < 16; k += 2 * off) acc[k] = fadd(acc[k], acc[k + off]);
                    const float dot = fadd(acc[0], __shfl_xor_sync(FULL, acc[0], 1));
                    pf.mark(3);
                    const float a = FAST ? act_fg_fast(fadd(pre_v, dot), gate) : act_fg(fadd(pre_v, dot), gate);
                    const float o = __shfl_xor_sync(FULL, a, 2);
                    const float z = gate ? fmul(o, a) : fmul(a, o);
                    if (zq == 0) {
                        zsb[zj] = z;
                        rb[LY::R_Z + zidx] = z;
                    } else {
                        v2_st_async(r_z + (uint32_t)b * (LY::ROWF * 4), z, r_zbar + (uint32_t)b * 8u);
                    }
                }
                pf.mark(4);
                v2_chain_sync();
                if (!has_next_layer && ct == 0) mbar_arrive(&fullb[b]);   // last layer: its helper (skip -> tail) is on the sample chain
                pf.mark(5);
                // 3. partial dense 1x1, output r = ct over the CTA's 32 gated channels, + residual -> layer l+1
                if (has_next_layer) {
                    const float4 *z4 = reinterpret_cast<const float4 *>(zsb);
                    const float xres = rb[LY::R_XRAW + ct];
                    float acc[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 zv = z4[i];
                        float a = 0.0f;
                        a = ffma(wdn[i].x, zv.x, a);
                        a = ffma(wdn[i].y, zv.y, a);
                        a = ffma(wdn[i].z, zv.z, a);
                        a = ffma(wdn[i].w, zv.w, a);
                        acc[i] = a;
                    }
#pragma unroll
                    for (int off = 1; off < 8; off <<= 1)
#pragma unroll
                        for (int k = 0; k < 8; k += 2 * off) acc[k] = fadd(acc[k], acc[k + off]);
                    const float v = (m == 0) ? fadd(fadd(xres, bdv), acc[0]) : acc[0];
                    if (out_dsmem) {
                        const uint32_t ro = (uint32_t)b * (LY::ROWF * 4), bo = (uint32_t)b * 8u;
#pragma unroll
                        for (int k = 0; k < 4; ++k) v2_st_async(r_x[k] + ro, v, r_xb[k] + bo);
                    } else {
                        ll_post(dx, v, seq);
                    }
                    // the helper streams 50 KB of shared memory per row: release it only after the dense has read its operands
                    __syncwarp();
                    if (ct == 0) mbar_arrive(&fullb[b]);
                }
                pf.stamp(11);
                pf.mark(6);
            }
        }
        // the last step of every row is drawn here: in the loop, step t-1 is drawn when step t is fed
        if (L0 && m == 0 && ct < 32) {
            for (int b = 0; b < N; ++b) {
                const int step = p.T_row[b] - 1;
                if (step < 0) continue;
                const float *u = p.noise + ((size_t)b * p.T + step) * (nr_mix + 1);
                float gum = 0.0f;
                if (ct < nr_mix) gum = ld_nc_f32(u + ct);
                const float logistic = ld_nc_f32(u + nr_mix);
                v2_sample_warp<SH>(p, mb, b, step, ct, true, ab, b2v, gum, logistic, pf);
            }
        }
        if (PROF && p.prof && ct == 0) {
            for (int i = 0; i < 7; ++i) p.prof[(size_t)cta * 16 + i] = pf.acc[PROF ? i : 0];
            p.prof[(size_t)cta * 16 + 14] = pf.acc[PROF ? 10 : 0];
            p.prof[(size_t)cta * 16 + 15] = pf.acc[PROF ? 11 : 0];
            p.prof[(size_t)cta * 16 + 13] = pf.acc[PROF ? 9 : 0];
            p.prof[(size_t)cta * 16 + 11] = pf.acc[PROF ? 8 : 0];
            p.prof[(size_t)cta * 16 + 12] = pf.acc[PROF ? 7 : 0];       // layer 0: polling the tail partials
        }
        };
        if (l == 0) chain_group(std::true_type{});
        else chain_group(std::false_type{});
    }
    // no CTA of the cluster leaves while a peer may still store into its shared memory
    __syncthreads();
    v2_cluster_sync();
}

// =================================================================================================================
// Tail CTA mt of kernel B: relu(total skip) -> conv1 slice (S/Mt columns) -> relu -> partial conv2 over that slice
// (wavenet/model.py:158-165).  conv1 from registers with lane = 16 consecutive k (two canonical 8-element chunks) and
// 4 columns per lane: 4 conflict-free LDS.128 per thread instead of 16 eight-address ones (512 -> 128 LSU cycles per
// row-step), partial sums reduced by a transposing butterfly.  Same canonical plan as tail_role_s (t_post1 = 64).
template <class SH>
__device__ void tail_role_v2(const WnParams &p, int mt)
{
    using Post1 = typename SH::Post1;
    using Post2 = typename SH::Post2;
    constexpr int S = SH::S, Sm = SH::Sm, M = SH::M, St = SH::St, O = SH::O, Mt = SH::Mt;
    static_assert(S == 512 && St == 32 && Post1::TPC == 8 && Post1::N4 == 16 && Post1::U == 8 && Post1::NPASS == 1, "v2 tail assumes the cfg-2 conv1 shape");
    static_assert(Post2::NPASS == 1 && Post2::XS == Post2::CH, "v2 tail assumes an unpadded conv2 input");
    constexpr int AS = 20;                                   // padded stride of a lane's 16-float chunk of relu(total)
    float *smem = g_smem;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int N = p.N, L = p.L;
    const float *gimg = p.tail_img + (size_t)mt * p.tail_img_floats;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.tail_smem_floats, &bar);
    const float *b1 = smem + p.off_b1;
    float *sc = smem + p.tail_smem_floats;
    float *as1 = sc, *c1s = sc + 32 * AS;                    // 640 + 32 floats
    for (int i = tid; i < 32 * AS + 32; i += WN_NT) sc[i] = 0.0f;
    __syncthreads();
    // conv1 weights: column 4w + q, k = 16*lane + 4i + kk  <->  packed slot col*8 + k/64, float4 index (k%64)/4
    float4 w1r[4][4];
    {
        const float4 *pk = reinterpret_cast<const float4 *>(smem + p.post1.off);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) w1r[q][i] = pk[(size_t)((lane & 3) * 4 + i) * WN_NT + (4 * w + q) * 8 + (lane >> 2)];
    }
    float4 w2r[Post2::N4];
    load_wreg<Post2::N4>(w2r, smem + p.post2.off);
    const float b1v = b1[4 * w + (lane & 3)];
    const float *xc2 = c1s + (tid % Post2::TPC) * Post2::XS;
    const int o2 = tid / Post2::TPC;
    const bool lead2 = (tid % Post2::TPC) == 0 && o2 < O;
    const int xa0 = (tid >> 4) * AS + (tid & 15), xa1 = ((tid + WN_NT) >> 4) * AS + (tid & 15);
    const size_t rowa = (size_t)L * M * Sm;
    const u64 *src0 = p.mb_acc + ((size_t)(L - 1) * M) * Sm + tid;
    const u64 *dst0 = p.mb_c2 + (size_t)mt * O + o2;
    V2Ab ab{p.status, 0};
    const MBox mb = make_mbox(p);
    Prof pf(p.prof ? p.prof + (size_t)(p.L * SH::M + mt) * 16 : nullptr);
    for (int t = 0; t < p.T; ++t) {
        if ((t & 15) == 0 && __syncthreads_or(ab.dead)) break;
        const unsigned seq = (unsigned)t + 1u;
        for (int b = 0; b < N; ++b) {
            if (t >= p.T_row[b]) continue;
            pf.start();
            MDst dc{nullptr, nullptr};
            if (lead2) dc = mb_dst(mb, dst0 + (size_t)b * Mt * O);
            {
                float q[4];
                v2_ll_wait_n(mb, src0 + b * rowa, (size_t)WN_NT, 2, seq, ab, q);
                as1[xa0] = relu32(q[0]);
                as1[xa1] = relu32(q[1]);
            }
            __syncthreads();
            pf.mark(0);
            pf.stamp(10);
            {
                const float4 *x4 = reinterpret_cast<const float4 *>(as1 + lane * AS);
                const float4 x0 = x4[0], x1 = x4[1], x2 = x4[2], x3 = x4[3];
                float acc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float a = 0.0f, c = 0.0f;
                    a = ffma(w1r[q][0].x, x0.x, a); a = ffma(w1r[q][0].y, x0.y, a); a = ffma(w1r[q][0].z, x0.z, a); a = ffma(w1r[q][0].w, x0.w, a);
                    a = ffma(w1r[q][1].x, x1.x, a); a = ffma(w1r[q][1].y, x1.y, a); a = ffma(w1r[q][1].z, x1.z, a); a = ffma(w1r[q][1].w, x1.w, a);
                    c = ffma(w1r[q][2].x, x2.x, c); c = ffma(w1r[q][2].y, x2.y, c); c = ffma(w1r[q][2].z, x2.z, c); c = ffma(w1r[q][2].w, x2.w, c);
                    c = ffma(w1r[q][3].x, x3.x, c); c = ffma(w1r[q][3].y, x3.y, c); c = ffma(w1r[q][3].z, x3.z, c); c = ffma(w1r[q][3].w, x3.w, c);
                    acc[q] = fadd(a, c);
                }
                // lanes: canonical offsets 2, 4 (transposing), then 8, 16, 32
                float dot = v2_reduce4(acc, lane);           // lanes xor 1, 2 (keep value lane & 3), xor 4
                dot = fadd(dot, __shfl_xor_sync(FULL, dot, 8));
                dot = fadd(dot, __shfl_xor_sync(FULL, dot, 16));
                if (lane < 4) c1s[4 * w + lane] = relu32(fadd(b1v, dot));
            }
            __syncthreads();
            pf.mark(1);
            {
                const float dot = butterfly<Post2::TPC>(dot_wreg<Post2::N4, Post2::U>(w2r, reinterpret_cast<const float4 *>(xc2)));
                if (lead2) ll_post(dc, dot, seq);
            }
            pf.stamp(11);
            pf.mark(2);
        }
    }
    pf.flush();
}

// Kernel A: a run of consecutive layers [p.layer_base, p.layer_end) in clusters of CS CTAs = CS / 4 layers each.  With a
// die map, clusters on die 0 claim their layer group from the front of the run and clusters on die 1 from the back, so the
// chain crosses the die boundary once on its way down instead of wherever the hardware put the clusters.
// Launch shapes (wn_api.cu): 15 clusters of 8 for the whole stack, or 7 clusters of 16 (layers 0..27) plus one cluster of 8
// (layers 28, 29) as two concurrent launches -- 16-CTA clusters halve the number of L2 hops but only 7 are co-resident.
template <class SH, bool FAST, bool PROF, int CS, bool SR = false>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(V2_NT, 1) wn_layers_kernel_v2(const __grid_constant__ WnParams p)
{
    __shared__ int s_crole;
    constexpr int LPC = CS / 4;
    const unsigned crank = v2_cluster_ctarank();
    if (crank == 0 && threadIdx.x == 0) {
        int role = (int)v2_cluster_id();
        if (p.sm_die != nullptr) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int n_clusters = (int)(gridDim.x / CS);
            role = (p.sm_die[smid] == 0) ? atomicAdd(p.status + p.claim_slot, 1) : n_clusters - 1 - atomicAdd(p.status + p.claim_slot + 1, 1);
        }
        s_crole = role;
    }
    v2_cluster_sync();
    int crole;
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(crole) : "r"(v2_mapa(smem_u32(&s_crole), 0u)) : "memory");
    const int l_local = (int)(crank >> 2), m = (int)(crank & 3u);
    const int l = p.layer_base + crole * LPC + l_local;
    if (l < p.layer_end) {
        layer_role_v2<SH, FAST, PROF, CS, SR>(p, l, m, l_local);
    } else {
        v2_cluster_sync();
        __syncthreads();
        v2_cluster_sync();
    }
}

// Kernel B: the tail CTAs, on the SMs the clusters leave free.
template <class SH>
__global__ void __launch_bounds__(WN_NT, 1) wn_tail_kernel_v2(const __grid_constant__ WnParams p)
{
    tail_role_v2<SH>(p, (int)blockIdx.x);
}
