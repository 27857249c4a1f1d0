// wn_kernel_ws.cuh -- warp-specialised layer CTA (used by the cfg-2 shape).
//
// The single-group layer role processes (row, step) items strictly in order: chain work of row b+1 waits
// for the off-chain work (and its mailbox waits) of row b, which puts ~8 k cycles per row-step on every
// CTA's loop and couples all CTAs' jitter (profiles/r01_phase_profile.md).  Here the CTA is split:
//
//   CHAIN  group, warps 4-7 (the SMSP arbiter favours the higher warp id): poll x -> filter/gate (current tap) -> tanh*sigmoid -> partial dense -> post.
//          Weights in registers (96 per thread).  A thread owns BOTH the filter and the gate column of one
//          gated channel for one quarter of K, so the activation needs no partner shuffle and the broadcast
//          x loads are half of what one-column-per-thread needs.
//   HELPER group, warps 0-3: dilation-ring push, sibling z gather, skip 1x1 + running skip sum,
//          pre-activations of the next step (dilated tap + lc), all streamed from shared memory.
//
// Hand-off per row through two mbarriers (phase parity = step parity): full[b] (chain -> helper: x and z of
// this step are in rowbuf[b]) and pre_rdy[b] (helper -> chain: pre[b] for the next step is written and
// rowbuf[b] may be overwritten).  Groups synchronise internally with named barriers 1 and 2.
//
// Arithmetic, evaluation plan and weight packing are IDENTICAL to layer_role_s: a helper/chain thread simply
// evaluates the packed-thread slots its new role maps to.  Included by wn_kernel.cu after wn_kernel_static.cuh.
#pragma once

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, unsigned parity)
{
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity, Abort &ab)
{
    unsigned spins = 0;
    long long t0 = 0;
    while (!mbar_try(bar, parity))
        if (((++spins) & 0xffu) == 0 && watchdog_check(ab, t0)) break;
}
// named barrier over the 128 threads of one group, OR-reducing the abort flag
__device__ __forceinline__ int group_sync_or(int id, int flag)
{
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "barrier.red.or.pred q, %2, 128, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(flag), "r"(id)
        : "memory");
    return r;
}

template <class SH>
__device__ void layer_role_ws(const WnParams &p, int l, int m)
{
    using Cur = typename SH::Cur;        // packed for 256 threads: col = idx / 4, chunk = idx % 4, 8 float4
    using Lc = typename SH::Lc;
    using Gc = typename SH::Gc;
    using Dense = typename SH::Dense;    // col = idx / 2, chunk = idx % 2, 4 float4
    using Skip = typename SH::Skip;      // col = idx / 2, chunk = idx % 2, 16 float4
    static_assert(Cur::TPC == 4 && Cur::N4 == 8 && Cur::NPASS == 1, "ws mapping assumes the cfg-2 fg shape");
    static_assert(Dense::TPC == 2 && Dense::N4 == 4 && Dense::NPASS == 1, "ws mapping assumes the cfg-2 dense shape");
    static_assert(Skip::TPC == 2 && Skip::NPASS == 1 && Lc::TPC == 4 && Gc::TPC == 4, "ws mapping assumes the cfg-2 shapes");
    constexpr int R = SH::R, M = SH::M, Dm = SH::Dm, Sm = SH::Sm, D = SH::D, ncol2 = 2 * SH::Dm, HALF = WN_NT / 2;
    static_assert(R == HALF && D == HALF && Sm == HALF && ncol2 * Cur::TPC == WN_NT, "ws mapping: 128-wide vectors");
    float *smem = g_smem;
    const int tid = threadIdx.x;
    const int N = p.N, L = p.L;
    const int cta = l * M + m;
    const float *gimg = p.layer_img + (size_t)cta * p.layer_img_floats;
    __shared__ uint64_t bar;
    __shared__ uint64_t mb_full[WN_MAX_BATCH], mb_pre[WN_MAX_BATCH];
    load_image_tma(smem, gimg, p.layer_smem_floats, &bar);

    const float *bfg = smem + p.off_bfg, *bd = smem + p.off_bd, *bs = smem + p.off_bs;
    float *sc = smem + p.layer_smem_floats;
    float *xs_cur = sc + p.ls.xs_cur, *xs_old = sc + p.ls.xs_old, *lcs = sc + p.ls.lcs;
    float *zs_dense = sc + p.ls.zs_dense, *zs_skip = sc + p.ls.zs_skip, *gvec = sc + p.ls.gvec;
    float *bfgN = sc + p.ls.bfgN, *pre = sc + p.ls.pre;
    float *rowbuf = sc + p.ls.total_floats;                  // [N][R + Dm]: x_l(t) and own z of the row in flight
    constexpr int RB = R + Dm;

    const int d = p.dil[l];
    const int nin = (l == 0) ? 1 : M;
    float *ring_cta = p.ring + p.ring_off[l] + (size_t)m * N * d * R;
    Abort ab{p.status, 0};
    const MBox mb = make_mbox(p);
    Prof pf(p.prof ? p.prof + (size_t)cta * 16 : nullptr);
    const size_t rowx = (size_t)L * M * R, rowz = (size_t)L * M * Dm, rowa = (size_t)L * M * Sm;

    for (int i = tid; i < p.ls.bfgN; i += WN_NT) sc[i] = 0.0f;
    for (int i = tid; i < N * RB; i += WN_NT) rowbuf[i] = 0.0f;
    if (tid == 0)
        for (int b = 0; b < N; ++b) { mbar_init(&mb_full[b], 1); mbar_init(&mb_pre[b], 1); }
    __syncthreads();

    // ---- prologue (all 256 threads, identical to layer_role_s): gc fold, pre-activations for t = 0 --------
    {
        const float *xc_old = xs_old + (tid % Cur::TPC) * Cur::XS;
        const float *xc_lc = lcs + (tid % Lc::TPC) * Lc::XS;
        const float *xc_gc = gvec + (tid % Gc::TPC) * Gc::XS;
        const float *w_old = smem + p.old.off + tid * 4, *w_lc = smem + p.lc.off + tid * 4, *w_gc = smem + p.gc.off + tid * 4;
        const int grp4 = tid / 4;
        const bool lead4 = (tid % 4) == 0;
        for (int b = 0; b < N; ++b) {
            if (SH::HAS_GC) {
                if (tid < SH::G) gvec[Gc::xpad(tid)] = __ldg(p.gc_table + (size_t)p.gc_id[b] * SH::G + tid);
                __syncthreads();
                matvec_s<Gc>(w_gc, xc_gc, grp4, lead4, ncol2, [&](int col, float dot) { bfgN[b * ncol2 + col] = fadd(bfg[col], dot); });
            } else {
                if (tid < ncol2) bfgN[b * ncol2 + tid] = bfg[tid];
            }
            __syncthreads();
            // old = zeros, lc = zeros (xs_old / lcs are zero-initialised)
            float *pre_b = pre + b * ncol2;
            const float *bias_b = bfgN + b * ncol2;
            matvec_s<Cur>(w_old, xc_old, grp4, lead4, ncol2, [&](int col, float dot) { pre_b[col] = fadd(bias_b[col], dot); });
            if (SH::HAS_LC) {
                __syncthreads();
                matvec_s<Lc>(w_lc, xc_lc, grp4, lead4, ncol2, [&](int col, float dot) { pre_b[col] = fadd(pre_b[col], dot); });
            }
            __syncthreads();
        }
        if (tid == 0)
            for (int b = 0; b < N; ++b) mbar_arrive(&mb_pre[b]);      // phase 0: pre for t = 0 is ready
        __syncthreads();
    }

    // the SMSP arbiter issues the highest warp id first: the latency-critical chain gets warps 4-7
    if (tid >= HALF) {
        const int tid = (int)threadIdx.x - HALF;             // chain-local thread id 0..127
        // =========================== CHAIN group ===========================================================
        const int j = tid >> 2, c4 = tid & 3;                 // gated channel, K quarter
        const int rp = tid >> 1, c2 = tid & 1;                // dense output pair, K half
        float4 wf[8], wg[8], wd0[4], wd1[4];
        {
            const float4 *pk = reinterpret_cast<const float4 *>(smem + p.cur.off);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                wf[i] = pk[(size_t)i * WN_NT + (2 * j) * 4 + c4];
                wg[i] = pk[(size_t)i * WN_NT + (2 * j + 1) * 4 + c4];
            }
            const float4 *pd = reinterpret_cast<const float4 *>(smem + p.dense.off);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                wd0[i] = pd[(size_t)i * WN_NT + (2 * rp) * 2 + c2];
                wd1[i] = pd[(size_t)i * WN_NT + (2 * rp + 1) * 2 + c2];
            }
        }
        const float4 *xc_cur = reinterpret_cast<const float4 *>(xs_cur + c4 * Cur::XS);
        const float4 *xc_dense = reinterpret_cast<const float4 *>(zs_dense + c2 * Dense::XS);
        const int xp_x = Cur::xpad(tid);
        const int xp_zd = Dense::xpad(j);
        const bool lead_fg = c4 == 0, lead_dn = c2 == 0;
        const float bd0 = bd[2 * rp], bd1 = bd[2 * rp + 1];
        const u64 *mbx_in = p.mb_x + ((size_t)l * M) * R + tid;
        const u64 *mbx_out = p.mb_x + ((size_t)(l + 1) * M + m) * R + 2 * rp;
        const u64 *mbz_out = p.mb_z + ((size_t)l * M + m) * Dm + j;

        for (int t = 0; t < p.T; ++t) {
            const unsigned seq = (unsigned)t + 1u;
            for (int b = 0; b < N; ++b) {
                if (t >= p.T_row[b]) continue;
                pf.start();
                float *rb = rowbuf + b * RB;
                const MDst dx0 = mb_dst(mb, mbx_out + b * rowx), dx1 = mb_dst(mb, mbx_out + b * rowx + 1);
                const MDst dz = mb_dst(mb, mbz_out + b * rowz);
                // pre[b] for this step is written and rowbuf[b] is free
                mbar_wait(&mb_pre[b], (unsigned)t & 1u, ab);
                // 1. layer input
                {
                    float q[4];
                    ll_wait_n(mb, mbx_in + b * rowx, (size_t)R, nin, seq, ab, q);
                    float v = q[0];
#pragma unroll
                    for (int i = 1; i < 4; ++i)
                        if (i < nin) v = fadd(v, q[i]);
                    xs_cur[xp_x] = v;
                    rb[tid] = v;
                }
                if (group_sync_or(1, ab.flag)) return;
                pf.mark(0);
                // 2. filter + gate columns of channel j over K quarter c4, from registers
                float z;
                {
                    float4 xv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) xv[i] = xc_cur[i];
                    float af[8], ag[8];
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        float a = 0.0f, g = 0.0f;
                        a = ffma(wf[s].x, xv[s].x, a); a = ffma(wf[s].y, xv[s].y, a); a = ffma(wf[s].z, xv[s].z, a); a = ffma(wf[s].w, xv[s].w, a);
                        g = ffma(wg[s].x, xv[s].x, g); g = ffma(wg[s].y, xv[s].y, g); g = ffma(wg[s].z, xv[s].z, g); g = ffma(wg[s].w, xv[s].w, g);
                        af[s] = a; ag[s] = g;
                    }
#pragma unroll
                    for (int off = 1; off < 8; off <<= 1)
#pragma unroll
                        for (int c = 0; c < 8; c += 2 * off) { af[c] = fadd(af[c], af[c + off]); ag[c] = fadd(ag[c], ag[c + off]); }
                    float f = butterfly<4>(af[0]), g = butterfly<4>(ag[0]);
                    f = fadd(pre[b * ncol2 + 2 * j], f);
                    g = fadd(pre[b * ncol2 + 2 * j + 1], g);
                    // the four lanes of a channel hold the same f and g: even lanes evaluate tanh(f), odd lanes
                    // sigmoid(g), concurrently; one shuffle brings the partner's factor
                    const bool odd = (c4 & 1) != 0;
                    const float a = act_fg(odd ? g : f, odd);
                    const float o = __shfl_xor_sync(FULL, a, 1);
                    z = odd ? fmul(o, a) : fmul(a, o);
                    if (lead_fg) {
                        zs_dense[xp_zd] = z;
                        rb[R + j] = z;
                        if (M > 1) ll_post(dz, z, seq);
                    }
                }
                if (group_sync_or(1, 0)) return;
                // the last layer has no dense: its helper (skip -> tail) is on the sample chain, release it now;
                // elsewhere the helper has a whole step of slack and would only fight the dense phase for the LSU
                if (l + 1 == L && tid == 0) mbar_arrive(&mb_full[b]);
                pf.mark(1);
                // 3. partial dense for outputs 2rp, 2rp+1 over K half c2
                if (l + 1 < L) {
                    float4 zv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) zv[i] = xc_dense[i];
                    float a0[4], a1[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        float a = 0.0f, g = 0.0f;
                        a = ffma(wd0[s].x, zv[s].x, a); a = ffma(wd0[s].y, zv[s].y, a); a = ffma(wd0[s].z, zv[s].z, a); a = ffma(wd0[s].w, zv[s].w, a);
                        g = ffma(wd1[s].x, zv[s].x, g); g = ffma(wd1[s].y, zv[s].y, g); g = ffma(wd1[s].z, zv[s].z, g); g = ffma(wd1[s].w, zv[s].w, g);
                        a0[s] = a; a1[s] = g;
                    }
#pragma unroll
                    for (int off = 1; off < 4; off <<= 1)
#pragma unroll
                        for (int c = 0; c < 4; c += 2 * off) { a0[c] = fadd(a0[c], a0[c + off]); a1[c] = fadd(a1[c], a1[c + off]); }
                    float d0 = butterfly<2>(a0[0]), d1 = butterfly<2>(a1[0]);
                    if (lead_dn) {
                        float v0 = (m == 0) ? fadd(fadd(rb[2 * rp], bd0), d0) : d0;
                        float v1 = (m == 0) ? fadd(fadd(rb[2 * rp + 1], bd1), d1) : d1;
                        ll_post(dx0, v0, seq);
                        ll_post(dx1, v1, seq);
                    }
                    if (tid == 0) mbar_arrive(&mb_full[b]);    // helper may start on this row
                }
                pf.mark(2);
            }
        }
        pf.flush(tid == 0);
    } else {
        // =========================== HELPER group ==========================================================
        const int ht = tid;                                     // 0..127; evaluates packed slots ht and ht + 128
        const int c4 = ht & 3, c2 = ht & 1;
        const float4 *xc_old = reinterpret_cast<const float4 *>(xs_old + c4 * Cur::XS);
        const float4 *xc_lc = reinterpret_cast<const float4 *>(lcs + c4 * Lc::XS);
        const float4 *xc_skip = reinterpret_cast<const float4 *>(zs_skip + c2 * Skip::XS);
        const float4 *w_old = reinterpret_cast<const float4 *>(smem + p.old.off) + ht;
        const float4 *w_lc = reinterpret_cast<const float4 *>(smem + p.lc.off) + ht;
        const float4 *w_skip = reinterpret_cast<const float4 *>(smem + p.skip.off) + ht;
        const int xp_x = Cur::xpad(ht);
        const int xp_lc = (SH::HAS_LC && ht < SH::C) ? Lc::xpad(ht) : 0;
        const int g_mm = ht / Dm, g_j = ht % Dm;
        const int xp_zs = Skip::xpad(ht);
        const int colA4 = ht >> 2, colB4 = (ht + HALF) >> 2;    // fg / lc columns of the two slots
        const int colA2 = ht >> 1, colB2 = (ht + HALF) >> 1;    // skip columns of the two slots
        const bool lead4 = c4 == 0, lead2 = c2 == 0;
        const float bsA = bs[colA2], bsB = bs[colB2];
        const u64 *mbz_in = p.mb_z + ((size_t)l * M + g_mm) * Dm + g_j;
        const u64 *mba_in = p.mb_acc + ((size_t)(l > 0 ? l - 1 : 0) * M + m) * Sm;
        const u64 *mba_out = p.mb_acc + ((size_t)l * M + m) * Sm;
        Prof hp(nullptr);

        for (int t = 0; t < p.T; ++t) {
            const unsigned seq = (unsigned)t + 1u;
            for (int b = 0; b < N; ++b) {
                if (t >= p.T_row[b]) continue;
                float *rb = rowbuf + b * RB;
                const bool has_next = (t + 1 < p.T_row[b]);
                const MDst da0 = mb_dst(mb, mba_out + b * rowa + colA2), da1 = mb_dst(mb, mba_out + b * rowa + colB2);
                // early loads for the next step's pre-activations (independent of this step's x unless d == 1)
                float oldv = 0.0f, lcv = 0.0f;
                if (has_next && d >= 2) oldv = __ldcg(ring_cta + ((size_t)b * d + ((t + 1) % d)) * R + ht);
                if (SH::HAS_LC && has_next && ht < SH::C) {
                    long idx = (long)t - p.lc_shift;
                    if (p.lc_up != nullptr && idx >= 0 && idx < p.t_lc) lcv = __ldg(p.lc_up + ((size_t)b * p.t_lc + idx) * SH::C + ht);
                }
                pin(oldv); pin(lcv);
                mbar_wait(&mb_full[b], (unsigned)t & 1u, ab);
                // ring push, sibling z gather
                const float xme = rb[ht];
                if (d >= 2) __stcg(ring_cta + ((size_t)b * d + (t % d)) * R + ht, xme);
                if (d == 1) oldv = xme;
                {
                    float z = (g_mm == m) ? rb[R + g_j] : ((M > 1) ? ll_wait(mb, mbz_in + b * rowz, seq, ab) : 0.0f);
                    zs_skip[xp_zs] = z;
                }
                xs_old[xp_x] = oldv;
                if (SH::HAS_LC && ht < SH::C) lcs[xp_lc] = lcv;
                if (group_sync_or(2, ab.flag)) return;
                // skip 1x1 for columns colA2 / colB2 (packed slots ht, ht + 128) + running skip sum
                {
                    float4 zv[Skip::N4];
#pragma unroll
                    for (int i = 0; i < Skip::N4; ++i) zv[i] = xc_skip[i];
                    float dA, dB;
                    {
                        float4 wv[Skip::N4];
#pragma unroll
                        for (int i = 0; i < Skip::N4; ++i) wv[i] = w_skip[(size_t)i * WN_NT];
                        dA = butterfly<2>(dot_wreg<Skip::N4, Skip::U>(wv, zv));
#pragma unroll
                        for (int i = 0; i < Skip::N4; ++i) wv[i] = w_skip[(size_t)i * WN_NT + HALF];
                        dB = butterfly<2>(dot_wreg<Skip::N4, Skip::U>(wv, zv));
                    }
                    if (lead2) {
                        float vA = fadd(bsA, dA), vB = fadd(bsB, dB);
                        if (l > 0) {
                            vA = fadd(ll_wait(mb, mba_in + b * rowa + colA2, seq, ab), vA);
                            vB = fadd(ll_wait(mb, mba_in + b * rowa + colB2, seq, ab), vB);
                        }
                        ll_post(da0, vA, seq);
                        ll_post(da1, vB, seq);
                    }
                }
                // pre-activations of the next step: bias(+gc) + W_old . x_l(t+1-d) + W_lc . lc(t)
                if (has_next) {
                    float4 xo[Cur::N4];
#pragma unroll
                    for (int i = 0; i < Cur::N4; ++i) xo[i] = xc_old[i];
                    float pA, pB;
                    {
                        float4 wv[Cur::N4];
#pragma unroll
                        for (int i = 0; i < Cur::N4; ++i) wv[i] = w_old[(size_t)i * WN_NT];
                        pA = butterfly<4>(dot_wreg<Cur::N4, Cur::U>(wv, xo));
#pragma unroll
                        for (int i = 0; i < Cur::N4; ++i) wv[i] = w_old[(size_t)i * WN_NT + HALF];
                        pB = butterfly<4>(dot_wreg<Cur::N4, Cur::U>(wv, xo));
                    }
                    pA = fadd(bfgN[b * ncol2 + colA4], pA);
                    pB = fadd(bfgN[b * ncol2 + colB4], pB);
                    if (SH::HAS_LC) {
                        float4 xl[Lc::N4], wv[Lc::N4];
#pragma unroll
                        for (int i = 0; i < Lc::N4; ++i) xl[i] = xc_lc[i];
#pragma unroll
                        for (int i = 0; i < Lc::N4; ++i) wv[i] = w_lc[(size_t)i * WN_NT];
                        pA = fadd(pA, butterfly<4>(dot_wreg<Lc::N4, Lc::U>(wv, xl)));
#pragma unroll
                        for (int i = 0; i < Lc::N4; ++i) wv[i] = w_lc[(size_t)i * WN_NT + HALF];
                        pB = fadd(pB, butterfly<4>(dot_wreg<Lc::N4, Lc::U>(wv, xl)));
                    }
                    if (lead4) { pre[b * ncol2 + colA4] = pA; pre[b * ncol2 + colB4] = pB; }
                }
                if (group_sync_or(2, ab.flag)) return;
                if (ht == 0) mbar_arrive(&mb_pre[b]);              // pre[b] for t+1 written, rowbuf[b] free
            }
        }
        (void)hp;
    }
}
