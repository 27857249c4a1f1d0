// wn_kernel_static.cuh -- compile-time specialised instantiations of the persistent kernel.
//
// Same algorithm, mailboxes, evaluation plan and arithmetic as the runtime-shaped roles in wn_kernel.cu;
// the differences are purely mechanical:
//   * every matrix shape (threads per column, float4 per thread, sub-chains, passes) is a template
//     constant, so the matvecs are straight-line code with all loads issued up-front;
//   * the weights on the sample-to-sample critical chain -- current-tap filter/gate columns and the
//     dense 1x1 slice of a layer CTA, conv1/conv2 slices of a tail CTA -- are loaded ONCE into registers
//     and stay there for the whole utterance (the register file is the largest on-chip memory: 256 KB/SM);
//     the off-chain matrices (dilated tap, lc, gc, skip) stream from shared memory.
// Included by wn_kernel.cu inside its anonymous namespace.
#pragma once

template <int TPC_, int N4_, int U_, int NPASS_>
struct MS {
    static constexpr int TPC = TPC_, N4 = N4_, U = U_, NPASS = NPASS_;
    static constexpr int CH = 4 * N4_;
    static constexpr int GPP = WN_NT / TPC_;
    static constexpr int XS = ((N4_ % 2) == 1) ? CH : CH + 4;     // padded chunk stride (make_mat in wn_api.cu)
    __host__ __device__ static constexpr int xpad(int k) { return (k / CH) * XS + (k % CH); }
    __host__ static bool matches(const WnMat &m)
    {
        return m.t == TPC && m.ch == CH && m.V == 4 && m.u == U && m.npass == NPASS && m.in_smem == 1 && m.xstride == XS;
    }
};

// BASELINE configs[1]: R = D = 128, S = 512, M = 4, Mt = 16, MoL-10, lc 80, gc 32
struct ShapeCfg2 {
    static constexpr bool SCALAR = true, HAS_LC = true, HAS_GC = true, WS = false;
    static constexpr int R = 128, D = 128, M = 4, Dm = 32, S = 512, Sm = 128, Mt = 16, St = 32, O = 30, C = 80, G = 32, IFW = 32, Q = 256;
    using Cur = MS<4, 8, 8, 1>;
    using Lc = MS<4, 5, 1, 1>;
    using Gc = MS<4, 2, 2, 1>;
    using Dense = MS<2, 4, 4, 1>;
    using Skip = MS<2, 16, 8, 1>;
    using Post1 = MS<8, 16, 8, 1>;
    using Post2 = MS<8, 1, 1, 1>;
    using Causal = MS<2, 4, 4, 1>;
};
// same shape with the warp-specialised layer CTA (wn_kernel_ws.cuh): ~4 % more latency per step but the
// off-chain work leaves the in-order loop, so it wins once >= 10 rows are in flight (profiles/r01_rows_sweep.md)
struct ShapeCfg2WS : ShapeCfg2 {
    static constexpr bool WS = true;
};
// BASELINE configs[0]: R = D = 32, S = 512, M = 1, Mt = 16, mu-law 256, unconditioned
struct ShapeCfg1 {
    static constexpr bool SCALAR = false, HAS_LC = false, HAS_GC = false, WS = false;
    static constexpr int R = 32, D = 32, M = 1, Dm = 32, S = 512, Sm = 512, Mt = 16, St = 32, O = 256, C = 0, G = 0, IFW = 32, Q = 256;
    using Cur = MS<4, 2, 2, 1>;
    using Lc = MS<1, 1, 1, 1>;
    using Gc = MS<1, 1, 1, 1>;
    using Dense = MS<8, 1, 1, 1>;
    using Skip = MS<1, 8, 8, 2>;
    using Post1 = MS<8, 16, 8, 1>;
    using Post2 = MS<1, 8, 8, 1>;
    using Causal = MS<1, 1, 1, 1>;
};
// the reference's hparams.py defaults: R = D = 32, S = 512, M = 1, Mt = 16, MoL-10, lc 80, gc 32
struct ShapeHparams {
    static constexpr bool SCALAR = true, HAS_LC = true, HAS_GC = true, WS = false;
    static constexpr int R = 32, D = 32, M = 1, Dm = 32, S = 512, Sm = 512, Mt = 16, St = 32, O = 30, C = 80, G = 32, IFW = 32, Q = 256;
    using Cur = MS<4, 2, 2, 1>;
    using Lc = MS<4, 5, 1, 1>;
    using Gc = MS<4, 2, 2, 1>;
    using Dense = MS<8, 1, 1, 1>;
    using Skip = MS<1, 8, 8, 2>;
    using Post1 = MS<8, 16, 8, 1>;
    using Post2 = MS<8, 1, 1, 1>;
    using Causal = MS<8, 1, 1, 1>;
};

template <class SH>
__host__ bool shape_matches(const WnParams &p)
{
    bool ok = p.R == SH::R && p.D == SH::D && p.M == SH::M && p.Dm == SH::Dm && p.S == SH::S && p.Sm == SH::Sm &&
              p.Mt == SH::Mt && p.St == SH::St && p.O == SH::O && p.C == SH::C && p.G == SH::G &&
              (p.scalar_input != 0) == SH::SCALAR && (!SH::SCALAR || p.ifw == SH::IFW) && (SH::SCALAR || p.Q == SH::Q);
    ok = ok && SH::Cur::matches(p.cur) && SH::Cur::matches(p.old) && SH::Dense::matches(p.dense) && SH::Skip::matches(p.skip);
    ok = ok && SH::Post1::matches(p.post1) && SH::Post2::matches(p.post2);
    if (SH::HAS_LC) ok = ok && SH::Lc::matches(p.lc);
    if (SH::HAS_GC) ok = ok && SH::Gc::matches(p.gc);
    if (SH::SCALAR) ok = ok && SH::Causal::matches(p.causal);
    return ok;
}

// ---------------------------------------------------------------------------------------------
// register-operand dot: the thread's N4 float4 of weights are already in registers
template <int N4, int U>
__device__ __forceinline__ float dot_wreg(const float4 (&wv)[N4], const float4 *__restrict__ x4)
{
    float4 xv[N4];
#pragma unroll
    for (int i = 0; i < N4; ++i) xv[i] = x4[i];
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

template <int TPC>
__device__ __forceinline__ float butterfly(float acc)
{
#pragma unroll
    for (int off = 1; off < TPC; off <<= 1) acc = fadd(acc, __shfl_xor_sync(FULL, acc, off));
    return acc;
}

// shared-memory matvec with a static shape; `w` = packed base + tid*4 floats, `xc` = this thread's chunk
template <class M, class F>
__device__ __forceinline__ void matvec_s(const float *__restrict__ w, const float *__restrict__ xc, int grp, bool lead, int ncols, F &&epi)
{
#pragma unroll
    for (int pass = 0; pass < M::NPASS; ++pass) {
        float acc = dot_regs<M::N4, M::U>(reinterpret_cast<const float4 *>(w) + (size_t)pass * M::N4 * WN_NT,
                                          reinterpret_cast<const float4 *>(xc));
        acc = butterfly<M::TPC>(acc);
        const int col = pass * M::GPP + grp;
        if (lead && col < ncols) epi(col, acc);
    }
}

template <int N4>
__device__ __forceinline__ void load_wreg(float4 (&wv)[N4], const float *packed_base)
{
    const float4 *w4 = reinterpret_cast<const float4 *>(packed_base) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < N4; ++i) wv[i] = w4[(size_t)i * WN_NT];
}

// =============================================================================================
template <class SH>
__device__ void layer_role_s(const WnParams &p, int l, int m)
{
    using Cur = typename SH::Cur;
    using Lc = typename SH::Lc;
    using Gc = typename SH::Gc;
    using Dense = typename SH::Dense;
    using Skip = typename SH::Skip;
    constexpr int R = SH::R, M = SH::M, Dm = SH::Dm, Sm = SH::Sm, D = SH::D, ncol2 = 2 * SH::Dm;
    float *smem = g_smem;
    const int tid = threadIdx.x;
    const int N = p.N, L = p.L;
    const int cta = l * M + m;
    const float *gimg = p.layer_img + (size_t)cta * p.layer_img_floats;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.layer_smem_floats, &bar);

    const float *bfg = smem + p.off_bfg, *bd = smem + p.off_bd, *bs = smem + p.off_bs;
    float *sc = smem + p.layer_smem_floats;
    float *xs_cur = sc + p.ls.xs_cur, *xs_old = sc + p.ls.xs_old, *lcs = sc + p.ls.lcs, *xraw = sc + p.ls.xraw;
    float *zs_dense = sc + p.ls.zs_dense, *zs_skip = sc + p.ls.zs_skip, *gvec = sc + p.ls.gvec;
    float *bfgN = sc + p.ls.bfgN, *pre = sc + p.ls.pre;

    const int d = p.dil[l];
    const int nin = (l == 0) ? 1 : M;
    float *ring_cta = p.ring + p.ring_off[l] + (size_t)m * N * d * R;
    Abort ab{p.status, 0};
    const MBox mb = make_mbox(p);
    Prof pf(p.prof ? p.prof + (size_t)cta * 16 : nullptr);

    // ---- chain weights -> registers (resident for the whole launch) ---------------------------------
    float4 wcur[Cur::N4], wdense[Dense::N4];
    load_wreg<Cur::N4>(wcur, smem + p.cur.off);
    load_wreg<Dense::N4>(wdense, smem + p.dense.off);

    // ---- per-thread constants ------------------------------------------------------------------------
    const float *xc_cur = xs_cur + (tid % Cur::TPC) * Cur::XS;
    const float *xc_old = xs_old + (tid % Cur::TPC) * Cur::XS;
    const float *xc_lc = lcs + (tid % Lc::TPC) * Lc::XS;
    const float *xc_gc = gvec + (tid % Gc::TPC) * Gc::XS;
    const float *xc_dense = zs_dense + (tid % Dense::TPC) * Dense::XS;
    const float *xc_skip = zs_skip + (tid % Skip::TPC) * Skip::XS;
    const float *w_old = smem + p.old.off + tid * 4, *w_lc = smem + p.lc.off + tid * 4, *w_gc = smem + p.gc.off + tid * 4;
    const float *w_skip = smem + p.skip.off + tid * 4;
    const int xp_x = (tid < R) ? Cur::xpad(tid) : 0;                 // cur and old share the layout
    const int xp_lc = (SH::HAS_LC && tid < SH::C) ? Lc::xpad(tid) : 0;
    const int g_mm = (tid < D) ? tid / Dm : 0, g_j = (tid < D) ? tid % Dm : 0;
    const int xp_zgather = (tid < D) ? Skip::xpad(tid) : 0;
    const int fg_grp = tid / Cur::TPC;
    const bool fg_lead = (tid % Cur::TPC) == 0;
    const bool fg_valid = fg_grp < ncol2;
    const bool fg_gate = (fg_grp & 1) != 0;
    const bool fg_store = fg_valid && fg_lead && !fg_gate;
    const int fg_j = fg_grp >> 1;
    const int xp_zd = fg_store ? Dense::xpad(fg_j) : 0;
    const int xp_zs = fg_store ? Skip::xpad(m * Dm + fg_j) : 0;
    const int dn_r = tid / Dense::TPC;                               // dense output column (single pass)
    const bool dn_lead = (tid % Dense::TPC) == 0 && dn_r < R;
    const float dn_bd = dn_lead ? bd[dn_r] : 0.0f;
    const int sk_grp = tid / Skip::TPC;
    const bool sk_lead = (tid % Skip::TPC) == 0;
    const int old_grp = fg_grp, lc_grp = tid / Lc::TPC, gc_grp = tid / Gc::TPC;
    const bool lc_lead = (tid % Lc::TPC) == 0, gc_lead = (tid % Gc::TPC) == 0;
    const size_t rowx = (size_t)L * M * R, rowz = (size_t)L * M * Dm, rowa = (size_t)L * M * Sm;
    const u64 *mbx_in = p.mb_x + ((size_t)l * M) * R + tid;
    u64 *mbx_out = p.mb_x + ((size_t)(l + 1) * M + m) * R + dn_r;
    u64 *mbz_out = p.mb_z + ((size_t)l * M + m) * Dm + fg_j;
    const u64 *mbz_in = p.mb_z + ((size_t)l * M + g_mm) * Dm + g_j;
    const u64 *mba_in = p.mb_acc + ((size_t)(l > 0 ? l - 1 : 0) * M + m) * Sm;
    u64 *mba_out = p.mb_acc + ((size_t)l * M + m) * Sm;

    for (int i = tid; i < p.ls.bfgN; i += WN_NT) sc[i] = 0.0f;      // all padded vectors
    __syncthreads();

    auto compute_pre = [&](int b, int tn) {
        if (tid < R) {
            float v;
            if (d == 1) v = (tn == 0) ? 0.0f : xraw[tid];
            else v = __ldcg(ring_cta + ((size_t)b * d + (tn % d)) * R + tid);
            xs_old[xp_x] = v;
        }
        if (SH::HAS_LC && tid < SH::C) {
            long idx = (long)tn - 1 - p.lc_shift;
            float v = 0.0f;
            if (p.lc_up != nullptr && idx >= 0 && idx < p.t_lc) v = __ldg(p.lc_up + ((size_t)b * p.t_lc + idx) * SH::C + tid);
            lcs[xp_lc] = v;
        }
        __syncthreads();
        float *pre_b = pre + b * ncol2;
        const float *bias_b = bfgN + b * ncol2;
        matvec_s<Cur>(w_old, xc_old, old_grp, fg_lead, ncol2, [&](int col, float dot) { pre_b[col] = fadd(bias_b[col], dot); });
        if (SH::HAS_LC) {
            __syncthreads();
            matvec_s<Lc>(w_lc, xc_lc, lc_grp, lc_lead, ncol2, [&](int col, float dot) { pre_b[col] = fadd(pre_b[col], dot); });
        }
        __syncthreads();
    };

    for (int b = 0; b < N; ++b) {
        if (SH::HAS_GC) {
            if (tid < SH::G) gvec[Gc::xpad(tid)] = __ldg(p.gc_table + (size_t)p.gc_id[b] * SH::G + tid);
            __syncthreads();
            matvec_s<Gc>(w_gc, xc_gc, gc_grp, gc_lead, ncol2, [&](int col, float dot) { bfgN[b * ncol2 + col] = fadd(bfg[col], dot); });
        } else {
            if (tid < ncol2) bfgN[b * ncol2 + tid] = bfg[tid];
        }
        __syncthreads();
        compute_pre(b, 0);
    }

    for (int t = 0; t < p.T; ++t) {
        const unsigned seq = (unsigned)t + 1u;
        for (int b = 0; b < N; ++b) {
            if (t >= p.T_row[b]) continue;
            pf.start();
            // destinations of this step's posts, resolved before the data exists
            const MDst dx = mb_dst(mb, mbx_out + b * rowx);
            const MDst dz = mb_dst(mb, mbz_out + b * rowz);
            // 1. layer input = sum of the partial residual outputs of layer l-1
            if (tid < R) {
                float q[4];
                ll_wait_n(mb, mbx_in + b * rowx, (size_t)R, nin, seq, ab, q);
                float v = q[0];
#pragma unroll
                for (int i = 1; i < 4; ++i)
                    if (i < nin) v = fadd(v, q[i]);
                xs_cur[xp_x] = v;
                xraw[tid] = v;
            }
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(0);
            // 2. current-tap filter/gate from registers + gated activation
            {
                float acc = butterfly<Cur::TPC>(dot_wreg<Cur::N4, Cur::U>(wcur, reinterpret_cast<const float4 *>(xc_cur)));
                float pv = fg_valid ? pre[b * ncol2 + fg_grp] : 0.0f;
                float a = act_fg(fadd(pv, acc), fg_gate);
                float other = __shfl_xor_sync(FULL, a, Cur::TPC);
                if (fg_store) {
                    float z = fmul(a, other);
                    zs_dense[xp_zd] = z;
                    zs_skip[xp_zs] = z;
                    if (M > 1) ll_post(dz, z, seq);
                }
            }
            __syncthreads();
            pf.mark(1);
            // 3. partial dense 1x1 from registers + residual -> mailbox of layer l+1
            if (l + 1 < L) {
                float dot = butterfly<Dense::TPC>(dot_wreg<Dense::N4, Dense::U>(wdense, reinterpret_cast<const float4 *>(xc_dense)));
                if (dn_lead) {
                    float v = (m == 0) ? fadd(fadd(xraw[dn_r], dn_bd), dot) : dot;
                    ll_post(dx, v, seq);
                }
            }
            pf.mark(2);
            // ---- off the critical chain ----
            if (d >= 2 && tid < R) __stcg(ring_cta + ((size_t)b * d + (t % d)) * R + tid, xraw[tid]);
            if (M > 1 && tid < D && g_mm != m) zs_skip[xp_zgather] = ll_wait(mb, mbz_in + b * rowz, seq, ab);
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(3);
            {
                const u64 *src = mba_in + b * rowa;
                u64 *dst = mba_out + b * rowa;
                matvec_s<Skip>(w_skip, xc_skip, sk_grp, sk_lead, Sm, [&](int c, float dot) {
                    float v = fadd(bs[c], dot);
                    if (l > 0) v = fadd(ll_wait(mb, src + c, seq, ab), v);
                    ll_post(mb, dst + c, v, seq);
                });
            }
            pf.mark(4);
            if (t + 1 < p.T_row[b]) compute_pre(b, t + 1);
            else __syncthreads();
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(5);
        }
    }
    pf.flush();
}

// =============================================================================================
template <class SH>
__device__ void tail_role_s(const WnParams &p, int mt)
{
    using Post1 = typename SH::Post1;
    using Post2 = typename SH::Post2;
    constexpr int S = SH::S, Sm = SH::Sm, M = SH::M, St = SH::St, O = SH::O, Mt = SH::Mt;
    float *smem = g_smem;
    const int tid = threadIdx.x;
    const int N = p.N, L = p.L;
    const float *gimg = p.tail_img + (size_t)mt * p.tail_img_floats;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.tail_smem_floats, &bar);
    const float *b1 = smem + p.off_b1;
    float *sc = smem + p.tail_smem_floats;
    float *as1 = sc + p.ts.as1, *c1s = sc + p.ts.c1s;
    for (int i = tid; i < p.ts.total_floats; i += WN_NT) sc[i] = 0.0f;
    __syncthreads();
    float4 w1r[Post1::N4], w2r[Post2::N4];
    static_assert(Post1::NPASS == 1 && Post2::NPASS == 1, "tail matvecs are single pass");
    load_wreg<Post1::N4>(w1r, smem + p.post1.off);
    load_wreg<Post2::N4>(w2r, smem + p.post2.off);
    const float *xc1 = as1 + (tid % Post1::TPC) * Post1::XS;
    const float *xc2 = c1s + (tid % Post2::TPC) * Post2::XS;
    const int c1 = tid / Post1::TPC;
    const bool lead1 = (tid % Post1::TPC) == 0 && c1 < St;
    const float b1v = lead1 ? b1[c1] : 0.0f;
    const int xp_c1 = lead1 ? Post2::xpad(c1) : 0;
    const int o2 = tid / Post2::TPC;
    const bool lead2 = (tid % Post2::TPC) == 0 && o2 < O;
    constexpr int PER = S / WN_NT;                          // acc words polled per thread
    static_assert(S % WN_NT == 0 && PER <= 4, "tail poll layout");
    int xp_a[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) xp_a[i] = Post1::xpad(tid + i * WN_NT);
    const size_t rowa = (size_t)L * M * Sm;
    const u64 *src0 = p.mb_acc + ((size_t)(L - 1) * M) * Sm + tid;
    u64 *dst0 = p.mb_c2 + (size_t)mt * O + o2;
    Abort ab{p.status, 0};
    const MBox mb = make_mbox(p);
    Prof pf(p.prof ? p.prof + (size_t)(p.L * SH::M + mt) * 16 : nullptr);
    for (int t = 0; t < p.T; ++t) {
        const unsigned seq = (unsigned)t + 1u;
        for (int b = 0; b < N; ++b) {
            if (t >= p.T_row[b]) continue;
            pf.start();
            {
                float q[4];
                ll_wait_n(mb, src0 + b * rowa, (size_t)WN_NT, PER, seq, ab, q);
#pragma unroll
                for (int i = 0; i < PER; ++i) as1[xp_a[i]] = relu32(q[i]);
            }
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(0);
            pf.stamp(10);
            {
                float dot = butterfly<Post1::TPC>(dot_wreg<Post1::N4, Post1::U>(w1r, reinterpret_cast<const float4 *>(xc1)));
                if (lead1) c1s[xp_c1] = relu32(fadd(b1v, dot));
            }
            __syncthreads();
            pf.mark(1);
            {
                float dot = butterfly<Post2::TPC>(dot_wreg<Post2::N4, Post2::U>(w2r, reinterpret_cast<const float4 *>(xc2)));
                if (lead2) ll_post(mb, dst0 + (size_t)b * Mt * O, dot, seq);
            }
            pf.stamp(11);
            __syncthreads();
            pf.mark(2);
        }
    }
    pf.flush();
}

// =============================================================================================
template <class SH>
__device__ void sampler_role_s(const WnParams &p)
{
    using Causal = typename SH::Causal;
    constexpr int R = SH::R, O = SH::O, Q = SH::Q, ifw = SH::IFW, Mt = SH::Mt, M = SH::M;
    constexpr int nr = SH::SCALAR ? O / 3 : 0;
    float *smem = g_smem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N;
    const float *gimg = p.samp_img;
    __shared__ uint64_t bar;
    load_image_tma(smem, gimg, p.samp_smem_floats, &bar);
    const float *b2 = smem + p.off_b2;
    float *sc = smem + p.samp_smem_floats;
    float *c2s = sc + p.ss.c2s, *cq = sc + p.ss.cq, *cqx = sc + p.ss.cqx;
    int *ids = reinterpret_cast<int *>(sc + p.ss.ids);
    double *cdf = reinterpret_cast<double *>(sc + p.ss.cdf);
    double *red = reinterpret_cast<double *>(sc + p.ss.red);
    float *misc = sc + p.ss.misc;
    Abort ab{p.status, 0};
    const MBox mb = make_mbox(p);
    Prof pf(p.prof ? p.prof + (size_t)(p.L * SH::M + SH::Mt) * 16 : nullptr);

    for (int i = tid; i < p.ss.total_floats; i += WN_NT) sc[i] = 0.0f;
    __syncthreads();
    for (int i = tid; i < 2 * N; i += WN_NT) ids[i] = -1;
    __syncthreads();
    float4 wcr[Causal::N4];
    if (SH::SCALAR) load_wreg<Causal::N4>(wcr, smem + p.causal.off);
    const float *xc_c = cqx + (tid % Causal::TPC) * Causal::XS;
    const int c_r = tid / Causal::TPC;
    const bool c_lead = (tid % Causal::TPC) == 0 && c_r < R;
    const int xp_cq = (SH::SCALAR && tid < ifw) ? Causal::xpad(tid) : 0;
    const size_t rowx = (size_t)p.L * M * R;
    const float b2v = (tid < O) ? b2[tid] : 0.0f;

    auto feed = [&](int b, float x_in, unsigned seq) {
        u64 *dst = p.mb_x + b * rowx;
        if (SH::SCALAR) {
            float v = 0.0f;
            if (tid < ifw) v = (tid < ifw - 1) ? cq[b * ifw + tid + 1] : x_in;
            __syncthreads();
            if (tid < ifw) { cq[b * ifw + tid] = v; cqx[xp_cq] = v; }
            __syncthreads();
            float dot = butterfly<Causal::TPC>(dot_wreg<Causal::N4, Causal::U>(wcr, reinterpret_cast<const float4 *>(xc_c)));
            if (c_lead) ll_post(mb, dst + c_r, dot, seq);
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
            float gum = 0.0f, logistic = 0.0f, next_forced = 0.0f;
            double u64v = 0.0;
            if (SH::SCALAR) {
                const float *u = (const float *)p.uniforms + ((size_t)b * p.T + t) * (nr + 1);
                if (warp == 0 && lane < nr) gum = wn::log32(-wn::log32(ld_nc_f32(u + lane)));
                if (tid == 0) { float u2 = ld_nc_f32(u + nr); logistic = fsub(wn::log32(u2), wn::log32(fsub(1.0f, u2))); }
            } else {
                u64v = ld_nc_f64((const double *)p.uniforms + (size_t)b * p.T + t);
            }
            const bool has_next = (t + 1 < p.T_row[b]);
            if (has_next && t + 1 < p.n_forced) next_forced = ld_nc_f32(p.forced + (size_t)b * p.n_forced + t + 1);
            pin(gum); pin(logistic); pin(next_forced); pin(u64v);

            if (tid < O) {
                float v = b2v;
                const u64 *src = p.mb_c2 + ((size_t)b * Mt) * O + tid;
#pragma unroll
                for (int m0 = 0; m0 < Mt; m0 += 4) {
                    float q[4];
                    ll_wait_n(mb, src + (size_t)m0 * O, (size_t)O, (Mt - m0 < 4) ? (Mt - m0) : 4, seq, ab, q);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (m0 + i < Mt) v = fadd(v, q[i]);
                }
                c2s[tid] = v;
                if (p.out_logits) p.out_logits[((size_t)b * p.T + t) * O + tid] = v;
            }
            if (__syncthreads_or(ab.flag)) return;
            pf.mark(0);
            pf.stamp(10);

            float sample;
            if (SH::SCALAR) {
                if (warp == 0) {
                    float g = (lane < nr) ? fsub(c2s[lane], gum) : __int_as_float(0xff800000);
                    int k = lane;
#pragma unroll
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
            pf.stamp(11);
            pf.mark(2);
        }
    }
    pf.flush();
}
