// wn_train_fused.cuh -- fused forward of ONE dilation layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA),
// for the flagship shape R = D = 128 (BASELINE configs[3]); any other shape runs the cuBLASLt path of wn_train.cu.
//
// Reference arithmetic (wavenet/model.py:66-101, train mode), per output row (n, tau >= off_l):
//     [f | g] = x[tau-d] . Wfg[0] + x[tau] . Wfg[1] + lc[tau-off_l] . Wlc + b + gc[n]          (K = 128 + 128 + C)
//     z = tanh(f) * sigmoid(g);   x_next[tau] = x[tau] + z . Wd + bd;   skip operand <- z for the last OW steps
//
// One CTA = one tile of 128 consecutive rows (absolute-time layout of wn_train_kernels.cuh), 2 CTAs per SM:
//   warp 0   TMA producer: six 64-wide K blocks -- (x[tau-d], x[tau], lc) x 2 -- of A (128 x 64, bf16) and of the
//            K-major transposed weights B (256 x 64), 128B-swizzled, through a 2-stage full/empty mbarrier ring; then
//            the dense weights Wd^T into the freed stage.
//   warp 1   allocates 256 TMEM columns; one lane issues tcgen05.mma (M=128, N=256, K=16) x 24 into TMEM columns
//            [0,256), commits to `acc1_full`; later 8 x (N=128) for the dense 1x1 into columns [0,128) -> `acc2_full`.
//   warps 2-9  epilogue: thread = (row = TMEM lane, half of the channels).  tcgen05.ld f/g -> + bias + gc -> tanh.approx / sigmoid -> tanh,
//            sigmoid to TS (for the backward pass), z to the skip operand and -- 128B-swizzled, as a K-major A operand --
//            to shared memory for the dense MMA; then acc2 + x[tau] + bd -> x_next.
// The fp32 pre-activations never exist in HBM: the cuBLASLt path moves 4.1 GB per layer for them, this kernel ~0.6 GB.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wntf {

typedef __nv_bfloat16 bf16;

constexpr int TILE_M = 128, KB = 64, NFG = 256, ND = 128, NKB = 6;       // rows per tile, K block, N of the two MMAs, K blocks
constexpr int A_BYTES = TILE_M * KB * 2, B_BYTES = NFG * KB * 2, STAGE_BYTES = A_BYTES + B_BYTES;   // 16 KB + 32 KB
constexpr int N_STAGES = 2;
constexpr int SMEM_BYTES = N_STAGES * STAGE_BYTES + 1024;                 // + slack for the 1024 B alignment
constexpr int EPI_WARPS = 8, EPI_THREADS = EPI_WARPS * 32;                  // two warps per TMEM lane quadrant, each takes half the columns
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int TMEM_COLS = 256;
constexpr int KTOT = 384;                                                // K of the transposed filter|gate weights (zero padded)

struct FusedArgs {
    int l, d, off, SL, OW, T0, LD, zs_col0, do_dense, has_lc, N;
    long M, x_row0;               // rows per layer buffer; first row of layer l inside the stacked X tensor
    const float *bias;            // (2D) fp32 [filter | gate] or null
    const float *gcb;             // (N, 2D) fp32 or null
    const float *bd;              // (R) fp32 or null
    const bf16 *Xl;               // layer input, (M, 128)
    bf16 *Xn;                     // layer output (M, 128), unused if !do_dense
    bf16 *TS;                     // (M, 256): tanh | sigmoid
    bf16 *Zs;                     // (N*OW, LD) concatenated skip operand
    unsigned *err;                // set to 1 if a barrier wait timed out
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU.  ~2 s at 2 GHz, then flag + trap (the launch fails with an error).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, unsigned *err) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            if (err) atomicExch(err, 1u);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, both operands K-major
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, one swizzle atom (64 bf16) along K:
// start address >> 4 in bits [0,14), stride byte offset (8 rows x 128 B = 1024) >> 4 in [32,46), version 1 in [46,48),
// layout type SWIZZLE_128B (2) in [61,64); the leading byte offset is unused for this layout (cute/arch/mma_sm100_desc.hpp).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (kind::f16): D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major (bits 15, 16 = 0),
// N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a lane that owns one row moves a whole 32-byte sector per instruction,
// which is what makes the thread = row (= TMEM lane) epilogues sector-efficient without a shared-memory transpose.
__device__ __forceinline__ void stg256(void *p, const uint32_t *v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                 "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ldg256(const void *p, uint32_t *v) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                 "=r"(v[6]), "=r"(v[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}

__global__ void __launch_bounds__(THREADS, 2)
layer_fwd_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_lc,
                       const __grid_constant__ CUtensorMap map_wfg, const __grid_constant__ CUtensorMap map_wd, const FusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[N_STAGES], empty_bar[N_STAGES], acc1_full, z_ready, wd_full, acc2_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[2][NFG];      // layer bias + global-condition vector of the (at most two) sentences a tile touches
    __shared__ __align__(16) float s_bd[ND];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row0 = (long)a.off + (long)blockIdx.x * TILE_M;      // first row of the tile inside the layer buffers

    if (threadIdx.x == 0) {
        for (int s = 0; s < N_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc1_full, 1);
        mbar_init(&z_ready, EPI_THREADS);
        mbar_init(&wd_full, 1);
        mbar_init(&acc2_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int nkb = a.has_lc ? NKB : 4;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % N_STAGES, ph = (kb / N_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1, a.err);
                uint8_t *sa = smem + s * STAGE_BYTES, *sb = sa + A_BYTES;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                if (kb < 2) tma_load_2d(sa, &map_x, &full_bar[s], kb * KB, (int)(a.x_row0 + row0 - a.d));
                else if (kb < 4) tma_load_2d(sa, &map_x, &full_bar[s], (kb - 2) * KB, (int)(a.x_row0 + row0));
                else tma_load_2d(sa, &map_lc, &full_bar[s], (kb - 4) * KB, (int)(row0 - a.off));
                tma_load_2d(sb, &map_wfg, &full_bar[s], kb * KB, a.l * NFG);
            }
            if (a.do_dense) {
                mbar_wait(&acc1_full, 0, a.err);                 // every filter|gate MMA has finished reading both stages
                uint8_t *sw = smem + 1 * STAGE_BYTES;
                mbar_expect_tx(&wd_full, 2 * ND * KB * 2);
                tma_load_2d(sw, &map_wd, &wd_full, 0, a.l * ND);
                tma_load_2d(sw + ND * KB * 2, &map_wd, &wd_full, KB, a.l * ND);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const int nkb = a.has_lc ? NKB : 4;
            const uint32_t idesc1 = instr_desc(TILE_M, NFG);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % N_STAGES, ph = (kb / N_STAGES) & 1;
                mbar_wait(&full_bar[s], ph, a.err);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
                const uint64_t da = smem_desc(sa), db = smem_desc(sb);
#pragma unroll
                for (int k = 0; k < KB / 16; ++k)
                    tc_mma(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc1, (kb | k) != 0 ? 1u : 0u);   // +32 B per K=16 step
                tc_commit(&empty_bar[s]);
            }
            tc_commit(&acc1_full);
            if (a.do_dense) {
                mbar_wait(&wd_full, 0, a.err);
                mbar_wait(&z_ready, 0, a.err);
                tc_fence_after();
                const uint32_t idesc2 = instr_desc(TILE_M, ND);
                const uint32_t sz = smem_u32(smem), sw = smem_u32(smem + STAGE_BYTES);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t da = smem_desc(sz + kb * A_BYTES), db = smem_desc(sw + kb * (ND * KB * 2));
#pragma unroll
                    for (int k = 0; k < KB / 16; ++k)
                        tc_mma(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc2, (kb | k) != 0 ? 1u : 0u);
                }
                tc_commit(&acc2_full);
            }
        }
    } else {
        // ===================== epilogue: thread = (row, column half) =====================
        const int q = warp & 3;                                  // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                        // which 64 of the 128 channels
        const int r = q * 32 + lane;                             // row inside the tile == TMEM lane
        const long row = row0 + r;
        const bool valid = row < a.M;
        const int n0 = (int)(row0 / a.T0);
        const int n = valid ? (int)(row / a.T0) : n0;
        const int tau = valid ? (int)(row - (long)n * a.T0) : 0;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        {   // stage bias (+ gc) for sentences n0 and n0+1, and the dense bias, while the MMAs run
            const int t = threadIdx.x - 64;
            const float b = a.bias ? a.bias[t] : 0.f;
            const int n1 = min(n0 + 1, a.N - 1);
            s_bias[0][t] = b + (a.gcb ? a.gcb[(size_t)n0 * NFG + t] : 0.f);
            s_bias[1][t] = b + (a.gcb ? a.gcb[(size_t)n1 * NFG + t] : 0.f);
            if (t < ND) s_bd[t] = a.bd ? a.bd[t] : 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        }
        const float *sb = s_bias[n - n0];
        const bool skip_row = valid && tau >= a.SL;
        bf16 *zs_row = skip_row ? a.Zs + ((size_t)n * a.OW + (tau - a.SL)) * a.LD + a.zs_col0 : nullptr;
        bf16 *ts_row = a.TS + (size_t)(valid ? row : 0) * NFG;

        mbar_wait(&acc1_full, 0, a.err);
        tc_fence_after();
#pragma unroll 1
        for (int jj = 0; jj < 4; ++jj) {
            const int c0 = (half * 4 + jj) * 16;
            float f[16], g[16];
            tc_ld16(tlane + c0, f);
            tc_ld16(tlane + 128 + c0, g);
            tc_ld_wait();
            uint32_t th_p[8], sg_p[8], z_p[8];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 bf4 = *reinterpret_cast<const float4 *>(sb + c0 + i), bg4 = *reinterpret_cast<const float4 *>(sb + 128 + c0 + i);
                const float fb[4] = {bf4.x, bf4.y, bf4.z, bf4.w}, gbv[4] = {bg4.x, bg4.y, bg4.z, bg4.w};
#pragma unroll
                for (int u = 0; u < 4; u += 2) {
                    const float t0 = tanh_fast(f[i + u] + fb[u]), t1 = tanh_fast(f[i + u + 1] + fb[u + 1]);
                    const float s0 = fmaf(0.5f, tanh_fast(0.5f * (g[i + u] + gbv[u])), 0.5f);
                    const float s1 = fmaf(0.5f, tanh_fast(0.5f * (g[i + u + 1] + gbv[u + 1])), 0.5f);
                    th_p[(i + u) >> 1] = pack2(t0, t1);
                    sg_p[(i + u) >> 1] = pack2(s0, s1);
                    z_p[(i + u) >> 1] = pack2(t0 * s0, t1 * s1);
                }
            }
            if (valid) {
                uint4 *pt = reinterpret_cast<uint4 *>(ts_row + c0), *ps = reinterpret_cast<uint4 *>(ts_row + 128 + c0);
                pt[0] = make_uint4(th_p[0], th_p[1], th_p[2], th_p[3]);
                pt[1] = make_uint4(th_p[4], th_p[5], th_p[6], th_p[7]);
                ps[0] = make_uint4(sg_p[0], sg_p[1], sg_p[2], sg_p[3]);
                ps[1] = make_uint4(sg_p[4], sg_p[5], sg_p[6], sg_p[7]);
                if (skip_row) {
                    uint4 *pz = reinterpret_cast<uint4 *>(zs_row + c0);
                    pz[0] = make_uint4(z_p[0], z_p[1], z_p[2], z_p[3]);
                    pz[1] = make_uint4(z_p[4], z_p[5], z_p[6], z_p[7]);
                }
            }
            if (a.do_dense) {
                // z as the K-major, 128B-swizzled A operand of the dense MMA: K block c0/64, 16-byte chunk index XOR (row & 7)
                uint8_t *zt = smem + (c0 >> 6) * A_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
                const int ch = (c0 & 63) >> 3;
                *reinterpret_cast<uint4 *>(zt + (((ch) ^ (r & 7)) << 4)) = make_uint4(z_p[0], z_p[1], z_p[2], z_p[3]);
                *reinterpret_cast<uint4 *>(zt + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(z_p[4], z_p[5], z_p[6], z_p[7]);
            }
        }
        if (a.do_dense) {
            tc_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of z -> visible to the tensor core
            mbar_arrive(&z_ready);
            const bf16 *x_row = a.Xl + (size_t)(valid ? row : 0) * ND;
            bf16 *xn_row = a.Xn + (size_t)(valid ? row : 0) * ND;
            // residual input of this thread's 64 channels: issued before the wait so the loads overlap the dense MMA
            uint4 xr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xr[i] = valid ? *reinterpret_cast<const uint4 *>(x_row + half * 64 + i * 8) : make_uint4(0, 0, 0, 0);
            mbar_wait(&acc2_full, 0, a.err);
            tc_fence_after();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c0 = (half * 4 + jj) * 16;
                float v[16];
                tc_ld16(tlane + c0, v);
                tc_ld_wait();
                const uint32_t xs[8] = {xr[2 * jj].x, xr[2 * jj].y, xr[2 * jj].z, xr[2 * jj].w, xr[2 * jj + 1].x, xr[2 * jj + 1].y, xr[2 * jj + 1].z, xr[2 * jj + 1].w};
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162 *>(&xs[i]);
                    const float v0 = v[2 * i] + __low2float(xb) + s_bd[c0 + 2 * i], v1 = v[2 * i + 1] + __high2float(xb) + s_bd[c0 + 2 * i + 1];
                    o[i] = pack2(v0, v1);
                }
                if (valid) {
                    uint4 *po = reinterpret_cast<uint4 *>(xn_row + c0);
                    po[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    po[1] = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent version of the forward kernel: one CTA per SM loops over tiles (tile = blockIdx.x + it*gridDim.x) with three
// decoupled pipelines, so that TMA, tensor core and epilogue of DIFFERENT tiles overlap instead of running back to back:
//   * a 3-stage shared-memory ring (48 KB per stage) that the producer keeps full ACROSS tile boundaries,
//   * two 256-column TMEM accumulators: the filter|gate MMAs of tile it+1 run while the epilogue drains tile it,
//   * the dense 1x1 of tile it is issued after the filter|gate MMAs of tile it+1 (its operand z comes from the epilogue) and
//     accumulates into columns [0,128) of tile it's own buffer, which the gate epilogue has finished reading.
// The dense weights Wd^T stay resident in shared memory for the whole launch; z has a dedicated 32 KB tile.
constexpr int PF_STAGES = 3;
constexpr int PF_OFF_WD = PF_STAGES * STAGE_BYTES;                    // 144 KB
constexpr int PF_OFF_Z = PF_OFF_WD + 2 * ND * KB * 2;                  // + 32 KB
constexpr int PF_SMEM_BYTES = PF_OFF_Z + 2 * A_BYTES + 1024;           // + 32 KB + alignment slack = 209 KB
constexpr int PF_TMEM_COLS = 512;
constexpr int PF_EPI_WARPS = 16, PF_EPI_THREADS = PF_EPI_WARPS * 32, PF_THREADS = 64 + PF_EPI_THREADS;   // four warps per TMEM lane quadrant, 32 channels each

// 18 warps: the register file is split per SM sub-partition (16 K registers, 5 warps on the fullest one) -> at most 96 per thread
__global__ void __launch_bounds__(PF_THREADS, 1)
layer_fwd_persistent_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_lc,
                            const __grid_constant__ CUtensorMap map_wfg, const __grid_constant__ CUtensorMap map_wd, const FusedArgs a, int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[PF_STAGES], empty_bar[PF_STAGES], acc1_full[2], acc2_full[2], tmem_empty[2], z_ready, wd_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[2][NFG];
    __shared__ __align__(16) float s_bd[ND];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = a.has_lc ? NKB : 4;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < PF_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc1_full[b], 1);
            mbar_init(&acc2_full[b], 1);
            mbar_init(&tmem_empty[b], PF_EPI_THREADS / 2);
        }
        mbar_init(&z_ready, PF_EPI_THREADS / 2);
        mbar_init(&wd_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)PF_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            if (a.do_dense) {
                mbar_expect_tx(&wd_full, 2 * ND * KB * 2);
                tma_load_2d(smem + PF_OFF_WD, &map_wd, &wd_full, 0, a.l * ND);
                tma_load_2d(smem + PF_OFF_WD + ND * KB * 2, &map_wd, &wd_full, KB, a.l * ND);
            }
            int g = 0;                                            // global K-block counter: the ring does not restart per tile
            for (int it = 0; it < my_tiles; ++it) {
                const long row0 = (long)a.off + ((long)blockIdx.x + (long)it * gridDim.x) * TILE_M;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % PF_STAGES, ph = (g / PF_STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1, a.err);
                    uint8_t *sa = smem + s * STAGE_BYTES, *sb = sa + A_BYTES;
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    if (kb < 2) tma_load_2d(sa, &map_x, &full_bar[s], kb * KB, (int)(a.x_row0 + row0 - a.d));
                    else if (kb < 4) tma_load_2d(sa, &map_x, &full_bar[s], (kb - 2) * KB, (int)(a.x_row0 + row0));
                    else tma_load_2d(sa, &map_lc, &full_bar[s], (kb - 4) * KB, (int)(row0 - a.off));
                    tma_load_2d(sb, &map_wfg, &full_bar[s], kb * KB, a.l * NFG);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc1 = instr_desc(TILE_M, NFG), idesc2 = instr_desc(TILE_M, ND);
            const uint32_t sz = smem_u32(smem + PF_OFF_Z), sw = smem_u32(smem + PF_OFF_WD);
            int g = 0;
            for (int it = 0; it <= my_tiles; ++it) {
                if (it < my_tiles) {
                    const int buf = it & 1;
                    mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1, a.err);        // epilogue has drained this accumulator (tile it-2)
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + (uint32_t)(buf * NFG);
                    for (int kb = 0; kb < nkb; ++kb, ++g) {
                        const int s = g % PF_STAGES, ph = (g / PF_STAGES) & 1;
                        mbar_wait(&full_bar[s], ph, a.err);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
                        const uint64_t da = smem_desc(sa), db = smem_desc(sb);
#pragma unroll
                        for (int k = 0; k < KB / 16; ++k) tc_mma(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc1, (kb | k) != 0 ? 1u : 0u);
                        tc_commit(&empty_bar[s]);
                    }
                    tc_commit(&acc1_full[buf]);
                }
                if (a.do_dense && it >= 1) {                      // dense 1x1 of the previous tile
                    const int pt = it - 1, pbuf = pt & 1;
                    if (pt == 0) mbar_wait(&wd_full, 0, a.err);
                    mbar_wait(&z_ready, pt & 1, a.err);
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + (uint32_t)(pbuf * NFG);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t da = smem_desc(sz + kb * A_BYTES), db = smem_desc(sw + kb * (ND * KB * 2));
#pragma unroll
                        for (int k = 0; k < KB / 16; ++k) tc_mma(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc2, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(&acc2_full[pbuf]);
                }
            }
        }
    } else {
        // ===================== epilogue: two warp groups working on consecutive tiles =====================
        // warps 2-9  GATE group: tile it: tcgen05.ld f/g -> gate -> tanh, sigmoid, skip z to HBM, z to shared memory (A of the dense MMA)
        // warps 10-17 RESIDUAL group: tile it: dense accumulator + x + bd -> x_next
        // so the gate epilogue of tile it+1 overlaps the dense MMA and the residual epilogue of tile it.  thread = (row, 64 channels).
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const bool gate_group = warp < 2 + PF_EPI_WARPS / 2;
        const int hb = ((warp - 2) >> 2) & 1;                    // which 64 of the 128 channels
        const int t = (threadIdx.x - 64) & (PF_EPI_THREADS / 2 - 1);   // index inside the group
        if (gate_group) {
            int staged_n0 = -1;
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;
                const long row0 = (long)a.off + ((long)blockIdx.x + (long)it * gridDim.x) * TILE_M;
                const long row = row0 + r;
                const bool valid = row < a.M;
                const int n0 = (int)(row0 / a.T0);
                const int n = valid ? (int)(row / a.T0) : n0;
                const int tau = valid ? (int)(row - (long)n * a.T0) : 0;
                const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NFG);
                if (n0 != staged_n0) {                           // uniform across the group: bias (+ gc) of sentences n0, n0+1
                    asm volatile("bar.sync 1, %0;" ::"n"(PF_EPI_THREADS / 2) : "memory");
                    const float b = a.bias ? a.bias[t] : 0.f;
                    const int n1 = min(n0 + 1, a.N - 1);
                    s_bias[0][t] = b + (a.gcb ? a.gcb[(size_t)n0 * NFG + t] : 0.f);
                    s_bias[1][t] = b + (a.gcb ? a.gcb[(size_t)n1 * NFG + t] : 0.f);
                    asm volatile("bar.sync 1, %0;" ::"n"(PF_EPI_THREADS / 2) : "memory");
                    staged_n0 = n0;
                }
                const float *sb = s_bias[n - n0];
                const bool skip_row = valid && tau >= a.SL;
                bf16 *zs_row = skip_row ? a.Zs + ((size_t)n * a.OW + (tau - a.SL)) * a.LD + a.zs_col0 : nullptr;
                bf16 *ts_row = a.TS + (size_t)(valid ? row : 0) * NFG;
                mbar_wait(&acc1_full[buf], (it >> 1) & 1, a.err);
                tc_fence_after();
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    const int cb = hb * 64 + hh * 32;
                    float f[32], g[32];
                    tc_ld32(tlane + cb, f);
                    tc_ld32(tlane + 128 + cb, g);
                    tc_ld_wait();
                    uint32_t th_p[16], sg_p[16], z_p[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 bf4 = *reinterpret_cast<const float4 *>(sb + cb + i), bg4 = *reinterpret_cast<const float4 *>(sb + 128 + cb + i);
                        const float fb[4] = {bf4.x, bf4.y, bf4.z, bf4.w}, gbv[4] = {bg4.x, bg4.y, bg4.z, bg4.w};
#pragma unroll
                        for (int u = 0; u < 4; u += 2) {
                            const float t0 = tanh_fast(f[i + u] + fb[u]), t1 = tanh_fast(f[i + u + 1] + fb[u + 1]);
                            const float s0 = fmaf(0.5f, tanh_fast(0.5f * (g[i + u] + gbv[u])), 0.5f);
                            const float s1 = fmaf(0.5f, tanh_fast(0.5f * (g[i + u + 1] + gbv[u + 1])), 0.5f);
                            th_p[(i + u) >> 1] = pack2(t0, t1);
                            sg_p[(i + u) >> 1] = pack2(s0, s1);
                            z_p[(i + u) >> 1] = pack2(t0 * s0, t1 * s1);
                        }
                    }
                    if (valid) {
                        stg256(ts_row + cb, th_p);
                        stg256(ts_row + cb + 16, th_p + 8);
                        stg256(ts_row + 128 + cb, sg_p);
                        stg256(ts_row + 128 + cb + 16, sg_p + 8);
                        if (skip_row) {
                            stg256(zs_row + cb, z_p);
                            stg256(zs_row + cb + 16, z_p + 8);
                        }
                    }
                    if (a.do_dense) {
                        // the single z tile is free once the dense MMA of the previous tile has completed (it is queued right
                        // behind this tile's filter|gate MMAs, so this wait is normally already satisfied)
                        if (hh == 0 && it >= 1) mbar_wait(&acc2_full[(it - 1) & 1], ((it - 1) >> 1) & 1, a.err);
                        uint8_t *zt = smem + PF_OFF_Z + (cb >> 6) * A_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
                        const int ch = (cb & 63) >> 3;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            *reinterpret_cast<uint4 *>(zt + (((ch + i) ^ (r & 7)) << 4)) = make_uint4(z_p[4 * i], z_p[4 * i + 1], z_p[4 * i + 2], z_p[4 * i + 3]);
                    }
                }
                tc_fence_before();
                if (a.do_dense) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&z_ready);
                } else {
                    mbar_arrive(&tmem_empty[buf]);
                }
            }
        } else if (a.do_dense) {
            if (t < ND) s_bd[t] = a.bd ? a.bd[t] : 0.f;
            asm volatile("bar.sync 2, %0;" ::"n"(PF_EPI_THREADS / 2) : "memory");
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;
                const long row0 = (long)a.off + ((long)blockIdx.x + (long)it * gridDim.x) * TILE_M;
                const long row = row0 + r;
                const bool valid = row < a.M;
                const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NFG);
                const bf16 *x_row = a.Xl + (size_t)(valid ? row : 0) * ND + hb * 64;
                bf16 *xn_row = a.Xn + (size_t)(valid ? row : 0) * ND + hb * 64;
                uint32_t xs[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) xs[i] = 0u;
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) ldg256(x_row + i * 16, xs + i * 8);
                }
                mbar_wait(&acc2_full[buf], (it >> 1) & 1, a.err);
                tc_fence_after();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int cb = hb * 64 + hh * 32;
                    float v[32];
                    tc_ld32(tlane + cb, v);
                    tc_ld_wait();
                    uint32_t o[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162 *>(&xs[hh * 16 + i]);
                        o[i] = pack2(v[2 * i] + __low2float(xb) + s_bd[cb + 2 * i], v[2 * i + 1] + __high2float(xb) + s_bd[cb + 2 * i + 1]);
                    }
                    if (valid) {
                        stg256(xn_row + hh * 32, o);
                        stg256(xn_row + hh * 32 + 16, o + 8);
                    }
                }
                tc_fence_before();
                mbar_arrive(&tmem_empty[buf]);                    // this accumulator may be overwritten by tile it+2
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)PF_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of one dilation layer, two tcgen05 kernels sharing one skeleton (tile = 128 rows, N = 128 accumulator columns,
// 3-stage TMA ring of 64-wide K blocks, 8 epilogue warps = (row, half of the channels)):
//
//  MODE_GATE  dz = dx_next . Wd^T (K = 128)  ->  + skip-path gradient dZs (last OW steps)  ->  gate backward with the saved
//             tanh / sigmoid  ->  dFG = [dz*sg*(1-th^2) | dz*th*sg*(1-sg)] (exact zeros for tau < off_l), z = th*sg
//             re-materialised for the dense weight gradient.  Replaces cast + GEMM + gate_bwd_kernel of the cuBLASLt path.
//  MODE_DX    dx[tau] = dx_next[tau] + dFG[tau] . Wfg[1]^T + dFG[tau+d] . Wfg[0]^T  as ONE K = 512 GEMM over two row offsets
//             of dFG, residual added in the epilogue, bf16 out.  Replaces two fp32 read-modify-write GEMMs.
// Rows run from s_l = off_l - d (the layer's input start) so that rows [s_l, off_l) are (re)written every step.
constexpr int BW_STAGES = 3, BW_A_BYTES = TILE_M * KB * 2, BW_B_BYTES = ND * KB * 2, BW_STAGE_BYTES = BW_A_BYTES + BW_B_BYTES;   // 16 + 16 KB
constexpr int BW_SMEM_BYTES = BW_STAGES * BW_STAGE_BYTES + 1024;
constexpr int BW_TMEM_COLS = 128;
constexpr int MODE_GATE = 0, MODE_DX = 1;

struct BwdArgs {
    int l, d, off, s, SL, OW, T0, LD, zs_col0, has_dense;
    long M;
    const bf16 *TS;      // (M, 256) tanh | sigmoid                       MODE_GATE
    const bf16 *dZs;     // (N*OW, LD) gradient of the skip operand        MODE_GATE
    bf16 *dFG;           // (M, 256) out                                   MODE_GATE
    bf16 *Z;             // (M, 128) out, may be null                      MODE_GATE
    const bf16 *dXin;    // (M, 128) gradient w.r.t. the layer output      MODE_DX (residual)
    bf16 *dXout;         // (M, 128) gradient w.r.t. the layer input       MODE_DX
    unsigned *err;
};

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2)
layer_bwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const BwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[BW_STAGES], empty_bar[BW_STAGES], acc_full;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row0 = (long)a.s + (long)blockIdx.x * TILE_M;
    const int nkb = MODE == MODE_GATE ? (a.has_dense ? 2 : 0) : 8;

    if (threadIdx.x == 0) {
        for (int st = 0; st < BW_STAGES; ++st) {
            mbar_init(&full_bar[st], 1);
            mbar_init(&empty_bar[st], 1);
        }
        mbar_init(&acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)BW_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % BW_STAGES, ph = (kb / BW_STAGES) & 1;
                mbar_wait(&empty_bar[st], ph ^ 1, a.err);
                uint8_t *sa = smem + st * BW_STAGE_BYTES, *sb = sa + BW_A_BYTES;
                mbar_expect_tx(&full_bar[st], BW_STAGE_BYTES);
                if (MODE == MODE_GATE) tma_load_2d(sa, &map_a, &full_bar[st], kb * KB, (int)row0);
                else tma_load_2d(sa, &map_a, &full_bar[st], (kb & 3) * KB, (int)(row0 + (kb >= 4 ? a.d : 0)));
                tma_load_2d(sb, &map_b, &full_bar[st], kb * KB, a.l * ND);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            const uint32_t idesc = instr_desc(TILE_M, ND);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % BW_STAGES, ph = (kb / BW_STAGES) & 1;
                mbar_wait(&full_bar[st], ph, a.err);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + st * BW_STAGE_BYTES), sb = sa + BW_A_BYTES;
                const uint64_t da = smem_desc(sa), db = smem_desc(sb);
#pragma unroll
                for (int k = 0; k < KB / 16; ++k) tc_mma(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                tc_commit(&empty_bar[st]);
            }
            tc_commit(&acc_full);
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        const long row = row0 + r;
        const bool valid = row < a.M;
        const int n = valid ? (int)(row / a.T0) : 0;
        const int tau = valid ? (int)(row - (long)n * a.T0) : 0;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        if (MODE == MODE_GATE) {
            const bool live = valid && tau >= a.off;                     // rows before the layer's first output carry exact zeros
            const bool skip_row = live && tau >= a.SL;
            const bf16 *ts_row = a.TS + (size_t)(valid ? row : 0) * NFG;
            const bf16 *dzs_row = skip_row ? a.dZs + ((size_t)n * a.OW + (tau - a.SL)) * a.LD + a.zs_col0 : nullptr;
            bf16 *dfg_row = a.dFG + (size_t)(valid ? row : 0) * NFG;
            bf16 *z_row = a.Z ? a.Z + (size_t)(valid ? row : 0) * ND : nullptr;
            if (nkb > 0) {
                mbar_wait(&acc_full, 0, a.err);
                tc_fence_after();
            }
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
                const int c0 = (half * 4 + jj) * 16;
                // saved activations / skip gradient of these 16 channels first, so the loads fly while TMEM is read
                uint4 th_r[2], sg_r[2], dz_r[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    th_r[i] = live ? *reinterpret_cast<const uint4 *>(ts_row + c0 + i * 8) : make_uint4(0, 0, 0, 0);
                    sg_r[i] = live ? *reinterpret_cast<const uint4 *>(ts_row + 128 + c0 + i * 8) : make_uint4(0, 0, 0, 0);
                    dz_r[i] = skip_row ? *reinterpret_cast<const uint4 *>(dzs_row + c0 + i * 8) : make_uint4(0, 0, 0, 0);
                }
                float v[16];
                if (nkb > 0) {
                    tc_ld16(tlane + c0, v);
                    tc_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                const uint32_t thw[8] = {th_r[0].x, th_r[0].y, th_r[0].z, th_r[0].w, th_r[1].x, th_r[1].y, th_r[1].z, th_r[1].w};
                const uint32_t sgw[8] = {sg_r[0].x, sg_r[0].y, sg_r[0].z, sg_r[0].w, sg_r[1].x, sg_r[1].y, sg_r[1].z, sg_r[1].w};
                const uint32_t dzw[8] = {dz_r[0].x, dz_r[0].y, dz_r[0].z, dz_r[0].w, dz_r[1].x, dz_r[1].y, dz_r[1].z, dz_r[1].w};
                uint32_t df_p[8], dg_p[8], z_p[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const __nv_bfloat162 tb = *reinterpret_cast<const __nv_bfloat162 *>(&thw[i]), sb2 = *reinterpret_cast<const __nv_bfloat162 *>(&sgw[i]),
                                         zb = *reinterpret_cast<const __nv_bfloat162 *>(&dzw[i]);
                    const float t0 = __low2float(tb), t1 = __high2float(tb), s0 = __low2float(sb2), s1 = __high2float(sb2);
                    const float d0 = live ? v[2 * i] + __low2float(zb) : 0.f, d1 = live ? v[2 * i + 1] + __high2float(zb) : 0.f;
                    df_p[i] = pack2(d0 * s0 * (1.f - t0 * t0), d1 * s1 * (1.f - t1 * t1));
                    dg_p[i] = pack2(d0 * t0 * s0 * (1.f - s0), d1 * t1 * s1 * (1.f - s1));
                    z_p[i] = pack2(t0 * s0, t1 * s1);
                }
                if (valid) {
                    uint4 *pf = reinterpret_cast<uint4 *>(dfg_row + c0), *pg = reinterpret_cast<uint4 *>(dfg_row + 128 + c0);
                    pf[0] = make_uint4(df_p[0], df_p[1], df_p[2], df_p[3]);
                    pf[1] = make_uint4(df_p[4], df_p[5], df_p[6], df_p[7]);
                    pg[0] = make_uint4(dg_p[0], dg_p[1], dg_p[2], dg_p[3]);
                    pg[1] = make_uint4(dg_p[4], dg_p[5], dg_p[6], dg_p[7]);
                    if (z_row) {
                        uint4 *pz = reinterpret_cast<uint4 *>(z_row + c0);
                        pz[0] = make_uint4(z_p[0], z_p[1], z_p[2], z_p[3]);
                        pz[1] = make_uint4(z_p[4], z_p[5], z_p[6], z_p[7]);
                    }
                }
            }
        } else {
            const bf16 *x_row = a.dXin + (size_t)(valid ? row : 0) * ND;
            bf16 *o_row = a.dXout + (size_t)(valid ? row : 0) * ND;
            uint4 xr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xr[i] = valid ? *reinterpret_cast<const uint4 *>(x_row + half * 64 + i * 8) : make_uint4(0, 0, 0, 0);
            mbar_wait(&acc_full, 0, a.err);
            tc_fence_after();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c0 = (half * 4 + jj) * 16;
                float v[16];
                tc_ld16(tlane + c0, v);
                tc_ld_wait();
                const uint32_t xs[8] = {xr[2 * jj].x, xr[2 * jj].y, xr[2 * jj].z, xr[2 * jj].w, xr[2 * jj + 1].x, xr[2 * jj + 1].y, xr[2 * jj + 1].z, xr[2 * jj + 1].w};
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162 *>(&xs[i]);
                    o[i] = pack2(v[2 * i] + __low2float(xb), v[2 * i + 1] + __high2float(xb));
                }
                if (valid) {
                    uint4 *po = reinterpret_cast<uint4 *>(o_row + c0);
                    po[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    po[1] = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BW_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent versions of the two backward kernels: one CTA per SM, the layer's B operand (Wd: 32 KB, or [Wfg1 | Wfg0]: 128 KB)
// is loaded ONCE and stays in shared memory, only A tiles stream through a 4-stage ring that runs ahead across tiles, two
// 128-column TMEM accumulators decouple the MMAs of tile it+1 from the epilogue of tile it, 16 epilogue warps = (row, 32 channels).
// Per 128-row tile the TMA traffic drops from 64 + 64 KB (MODE_GATE) / 256 + 128 KB... to the A operand alone.
constexpr int PB_STAGES = 4;
constexpr int PB_OFF_B = PB_STAGES * BW_A_BYTES;                       // 64 KB of A stages, then the resident B (up to 8 K blocks x 16 KB)
constexpr int PB_WB_ROW = 80, PB_WB_BYTES = 32 * PB_WB_ROW;            // per-warp staging buffer: 32 rows x (64 B + 16 B pad)
template <int MODE> struct PbLayout {
    static constexpr int NB = MODE == MODE_GATE ? 2 : 8;                // resident K blocks of B
    static constexpr int OFF_W = PB_OFF_B + NB * BW_B_BYTES;            // MODE_GATE: 3 staging buffers (th, sg, dz) per epilogue warp
    static constexpr int SMEM = OFF_W + (MODE == MODE_GATE ? PF_EPI_WARPS * 3 * PB_WB_BYTES : 0) + 1024;   // 217 KB / 193 KB
};
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
constexpr int PB_TMEM_COLS = 256;

template <int MODE>
__global__ void __launch_bounds__(PF_THREADS, 1)
layer_bwd_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const BwdArgs a, int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[PB_STAGES], empty_bar[PB_STAGES], acc_full[2], tmem_empty[2], b_full;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = MODE == MODE_GATE ? (a.has_dense ? 2 : 0) : 8;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        for (int st = 0; st < PB_STAGES; ++st) {
            mbar_init(&full_bar[st], 1);
            mbar_init(&empty_bar[st], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&tmem_empty[b], PF_EPI_THREADS);
        }
        mbar_init(&b_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)PB_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0 && nkb > 0) {
            mbar_expect_tx(&b_full, nkb * BW_B_BYTES);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smem + PB_OFF_B + kb * BW_B_BYTES, &map_b, &b_full, kb * KB, a.l * ND);   // nkb <= PbLayout<MODE>::NB
            int g = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const long row0 = (long)a.s + ((long)blockIdx.x + (long)it * gridDim.x) * TILE_M;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int st = g % PB_STAGES, ph = (g / PB_STAGES) & 1;
                    mbar_wait(&empty_bar[st], ph ^ 1, a.err);
                    mbar_expect_tx(&full_bar[st], BW_A_BYTES);
                    if (MODE == MODE_GATE) tma_load_2d(smem + st * BW_A_BYTES, &map_a, &full_bar[st], kb * KB, (int)row0);
                    else tma_load_2d(smem + st * BW_A_BYTES, &map_a, &full_bar[st], (kb & 3) * KB, (int)(row0 + (kb >= 4 ? a.d : 0)));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            const uint32_t idesc = instr_desc(TILE_M, ND);
            mbar_wait(&b_full, 0, a.err);
            int g = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;
                mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1, a.err);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * ND);
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int st = g % PB_STAGES, ph = (g / PB_STAGES) & 1;
                    mbar_wait(&full_bar[st], ph, a.err);
                    tc_fence_after();
                    const uint64_t da = smem_desc(smem_u32(smem + st * BW_A_BYTES)), db = smem_desc(smem_u32(smem + PB_OFF_B + kb * BW_B_BYTES));
#pragma unroll
                    for (int k = 0; k < KB / 16; ++k) tc_mma(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&empty_bar[st]);
                }
                tc_commit(&acc_full[buf]);
            }
        }
    } else {
        const int q = warp & 3, qd = (warp - 2) >> 2;
        const int r = q * 32 + lane, cb = qd * 32;
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            const long row0 = (long)a.s + ((long)blockIdx.x + (long)it * gridDim.x) * TILE_M;
            const long row = row0 + r;
            const bool valid = row < a.M;
            const int n = valid ? (int)(row / a.T0) : 0;
            const int tau = valid ? (int)(row - (long)n * a.T0) : 0;
            const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * ND);
            if (MODE == MODE_GATE) {
                // Thread = row is what TMEM gives, but 32 lanes x 16 B at a 512 B stride costs 32 half-used sectors per request.
                // So the warp's 32 rows x 32 channels of tanh / sigmoid / skip gradient come in through coalesced cp.async
                // (4 lanes per row, 64 B contiguous) into a per-warp staging buffer, and dFG / z leave the same way.
                const bool live = valid && tau >= a.off;
                uint8_t *wb = smem + PbLayout<MODE>::OFF_W + (warp - 2) * 3 * PB_WB_BYTES;
                const long wrow0 = row0 + q * 32;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int id = k * 32 + lane, rw = id >> 2, part = id & 3;
                    const long rg = wrow0 + rw;
                    const int nr = (int)(rg / a.T0), tr = (int)(rg - (long)nr * a.T0);
                    const bool lv = rg < a.M && tr >= a.off;
                    uint8_t *d0 = wb + rw * PB_WB_ROW + part * 16;
                    if (lv) {
                        const bf16 *src = a.TS + (size_t)rg * NFG + cb + part * 8;
                        cp_async16(d0, src);
                        cp_async16(d0 + PB_WB_BYTES, src + 128);
                    } else {
                        *reinterpret_cast<uint4 *>(d0) = make_uint4(0, 0, 0, 0);
                        *reinterpret_cast<uint4 *>(d0 + PB_WB_BYTES) = make_uint4(0, 0, 0, 0);
                    }
                    if (lv && tr >= a.SL) cp_async16(d0 + 2 * PB_WB_BYTES, a.dZs + ((size_t)nr * a.OW + (tr - a.SL)) * a.LD + a.zs_col0 + cb + part * 8);
                    else *reinterpret_cast<uint4 *>(d0 + 2 * PB_WB_BYTES) = make_uint4(0, 0, 0, 0);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                float v[32];
                if (nkb > 0) {
                    mbar_wait(&acc_full[buf], (it >> 1) & 1, a.err);
                    tc_fence_after();
                    tc_ld32(tlane + cb, v);
                    tc_ld_wait();
                    tc_fence_before();
                    mbar_arrive(&tmem_empty[buf]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                uint8_t *mine = wb + lane * PB_WB_ROW;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {                 // 16 channels at a time keeps the live set under 96 registers
                    uint4 th_r[2], sg_r[2], dz_r[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        th_r[i] = *reinterpret_cast<const uint4 *>(mine + (hh * 2 + i) * 16);
                        sg_r[i] = *reinterpret_cast<const uint4 *>(mine + PB_WB_BYTES + (hh * 2 + i) * 16);
                        dz_r[i] = *reinterpret_cast<const uint4 *>(mine + 2 * PB_WB_BYTES + (hh * 2 + i) * 16);
                    }
                    const uint32_t *thw = reinterpret_cast<const uint32_t *>(th_r), *sgw = reinterpret_cast<const uint32_t *>(sg_r),
                                   *dzw = reinterpret_cast<const uint32_t *>(dz_r);
                    uint32_t df_p[8], dg_p[8], z_p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __nv_bfloat162 tb = *reinterpret_cast<const __nv_bfloat162 *>(&thw[i]), sb2 = *reinterpret_cast<const __nv_bfloat162 *>(&sgw[i]),
                                             zb = *reinterpret_cast<const __nv_bfloat162 *>(&dzw[i]);
                        const float t0 = __low2float(tb), t1 = __high2float(tb), s0 = __low2float(sb2), s1 = __high2float(sb2);
                        const float d0 = live ? v[hh * 16 + 2 * i] + __low2float(zb) : 0.f, d1 = live ? v[hh * 16 + 2 * i + 1] + __high2float(zb) : 0.f;
                        df_p[i] = pack2(d0 * s0 * (1.f - t0 * t0), d1 * s1 * (1.f - t1 * t1));
                        dg_p[i] = pack2(d0 * t0 * s0 * (1.f - s0), d1 * t1 * s1 * (1.f - s1));
                        z_p[i] = pack2(t0 * s0, t1 * s1);
                    }
                    // results overwrite this lane's own row of the three buffers
                    *reinterpret_cast<uint4 *>(mine + (hh * 2) * 16) = make_uint4(df_p[0], df_p[1], df_p[2], df_p[3]);
                    *reinterpret_cast<uint4 *>(mine + (hh * 2 + 1) * 16) = make_uint4(df_p[4], df_p[5], df_p[6], df_p[7]);
                    *reinterpret_cast<uint4 *>(mine + PB_WB_BYTES + (hh * 2) * 16) = make_uint4(dg_p[0], dg_p[1], dg_p[2], dg_p[3]);
                    *reinterpret_cast<uint4 *>(mine + PB_WB_BYTES + (hh * 2 + 1) * 16) = make_uint4(dg_p[4], dg_p[5], dg_p[6], dg_p[7]);
                    *reinterpret_cast<uint4 *>(mine + 2 * PB_WB_BYTES + (hh * 2) * 16) = make_uint4(z_p[0], z_p[1], z_p[2], z_p[3]);
                    *reinterpret_cast<uint4 *>(mine + 2 * PB_WB_BYTES + (hh * 2 + 1) * 16) = make_uint4(z_p[4], z_p[5], z_p[6], z_p[7]);
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int id = k * 32 + lane, rw = id >> 2, part = id & 3;
                    const long rg = wrow0 + rw;
                    if (rg < a.M) {
                        const uint8_t *s0 = wb + rw * PB_WB_ROW + part * 16;
                        bf16 *dst = a.dFG + (size_t)rg * NFG + cb + part * 8;
                        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(s0);
                        *reinterpret_cast<uint4 *>(dst + 128) = *reinterpret_cast<const uint4 *>(s0 + PB_WB_BYTES);
                        if (a.Z) *reinterpret_cast<uint4 *>(a.Z + (size_t)rg * ND + cb + part * 8) = *reinterpret_cast<const uint4 *>(s0 + 2 * PB_WB_BYTES);
                    }
                }
                __syncwarp();
            } else {
                const bf16 *x_row = a.dXin + (size_t)(valid ? row : 0) * ND;
                bf16 *o_row = a.dXout + (size_t)(valid ? row : 0) * ND;
                uint32_t xs[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) xs[i] = 0u;
                if (valid) {
                    ldg256(x_row + cb, xs);
                    ldg256(x_row + cb + 16, xs + 8);
                }
                mbar_wait(&acc_full[buf], (it >> 1) & 1, a.err);
                tc_fence_after();
                float v[32];
                tc_ld32(tlane + cb, v);
                tc_ld_wait();
                tc_fence_before();
                mbar_arrive(&tmem_empty[buf]);
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162 *>(&xs[i]);
                    o[i] = pack2(v[2 * i] + __low2float(xb), v[2 * i + 1] + __high2float(xb));
                }
                if (valid) {
                    stg256(o_row + cb, o);
                    stg256(o_row + cb + 16, o + 8);
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)PB_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradients of one dilation layer as split-K tcgen05 GEMMs over the ROW dimension (K = all 478 k rows):
//     dW[m][n] = sum_rows A[row][m] * B[row][n]
// Both operands are consumed exactly as they lie in HBM (row-major activations), i.e. MN-major for the tensor core: a TMA box
// of 64 rows x 64 channels (128 B per row, 128B swizzle) is one swizzle atom column of an MN-major operand -- 8-row groups
// 1024 B apart along K (stride byte offset), 64-channel groups one box (8 KB) apart along M/N (leading byte offset).
//   MODE_WFG  acc0 (128 x 256) = X[row-d]^T . dFG[row]   (tap 0),   acc1 (128 x 256) = X[row]^T . dFG[row]   (tap 1)
//   MODE_WLD  acc0 (128 x 256) = LC[row-off]^T . dFG[row] (rows >= C are zero: TMA out-of-bounds fill),  acc1 (128 x 128) = Z[row]^T . dXn[row]
// dFG is read ONCE for both accumulators (cuBLASLt reads it once per GEMM).  Each CTA owns every gridDim-th 64-row block, keeps
// its partial sums in TMEM (512 columns) and adds them to the fp32 gradient buffer with 16-byte vector atomics at the end.
constexpr int WG_KROWS = 64;                                           // rows per K block
constexpr int WG_BOX = WG_KROWS * 128;                                  // one 64 x 64 bf16 box = 8 KB
constexpr int MODE_WFG = 0, MODE_WLD = 1;
template <int MODE> struct WgLayout {
    static constexpr int A0 = 0, A1 = 2 * WG_BOX, B0 = 4 * WG_BOX;      // A0, A1: 128 channels = 2 boxes; B0 (dFG): 256 = 4 boxes
    static constexpr int B1 = 8 * WG_BOX;                               // MODE_WLD: dXn, 2 boxes
    static constexpr int STAGE = MODE == MODE_WFG ? 8 * WG_BOX : 10 * WG_BOX;   // 64 KB / 80 KB
    static constexpr int STAGES = MODE == MODE_WFG ? 3 : 2;
    static constexpr int SMEM = STAGES * STAGE + 1024;
};

struct WgArgs {
    int l, d, off, n_kblocks;
    long M, x_row0;            // rows per layer buffer; first row of layer l in the stacked X tensor
    float *dW0, *dW1;          // MODE_WFG: tap 0 / tap 1 (128 x 256 each);  MODE_WLD: dWlc (C x 256) / dWd (128 x 128)
    int rows0;                 // valid rows of dW0 (MODE_WLD: C), 128 otherwise
    unsigned *err;
};

// MN-major operand descriptor (see the comment above): LBO = 8 KB between 64-channel groups, SBO = 1 KB between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((uint32_t)WG_BOX >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t instr_desc_mn(int M, int N) { return instr_desc(M, N) | (1u << 15) | (1u << 16); }

template <int MODE>
__global__ void __launch_bounds__(PF_THREADS, 1)
layer_wgrad_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b0,
                   const __grid_constant__ CUtensorMap map_b1, const WgArgs a) {
    typedef WgLayout<MODE> LY;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[LY::STAGES], empty_bar[LY::STAGES], acc_full;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_blocks = ((int)blockIdx.x < a.n_kblocks) ? (a.n_kblocks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        for (int st = 0; st < LY::STAGES; ++st) {
            mbar_init(&full_bar[st], 1);
            mbar_init(&empty_bar[st], 1);
        }
        mbar_init(&acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < my_blocks; ++i) {
                const int st = i % LY::STAGES, ph = (i / LY::STAGES) & 1;
                const long row = (long)a.off + ((long)blockIdx.x + (long)i * gridDim.x) * WG_KROWS;      // rows past M read as zero
                mbar_wait(&empty_bar[st], ph ^ 1, a.err);
                uint8_t *sp = smem + st * LY::STAGE;
                mbar_expect_tx(&full_bar[st], LY::STAGE);
                if (MODE == MODE_WFG) {
                    for (int h = 0; h < 2; ++h) {
                        tma_load_2d(sp + LY::A0 + h * WG_BOX, &map_a0, &full_bar[st], h * 64, (int)(a.x_row0 + row - a.d));
                        tma_load_2d(sp + LY::A1 + h * WG_BOX, &map_a0, &full_bar[st], h * 64, (int)(a.x_row0 + row));
                    }
                } else {
                    for (int h = 0; h < 2; ++h) {
                        tma_load_2d(sp + LY::A0 + h * WG_BOX, &map_a0, &full_bar[st], h * 64, (int)(row - a.off));   // LC
                        tma_load_2d(sp + LY::A1 + h * WG_BOX, &map_a1, &full_bar[st], h * 64, (int)row);             // Z
                        tma_load_2d(sp + LY::B1 + h * WG_BOX, &map_b1, &full_bar[st], h * 64, (int)row);             // dXn
                    }
                }
                for (int h = 0; h < 4; ++h) tma_load_2d(sp + LY::B0 + h * WG_BOX, &map_b0, &full_bar[st], h * 64, (int)row);   // dFG
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id256 = instr_desc_mn(128, 256), id128 = instr_desc_mn(128, 128);
            for (int i = 0; i < my_blocks; ++i) {
                const int st = i % LY::STAGES, ph = (i / LY::STAGES) & 1;
                mbar_wait(&full_bar[st], ph, a.err);
                tc_fence_after();
                const uint32_t sp = smem_u32(smem + st * LY::STAGE);
#pragma unroll
                for (int k = 0; k < WG_KROWS / 16; ++k) {          // 16 rows = 2048 B further down the box
                    const uint64_t koff = (uint64_t)((k * 16 * 128) >> 4);
                    const uint64_t da0 = smem_desc_mn(sp + LY::A0) + koff, da1 = smem_desc_mn(sp + LY::A1) + koff, db0 = smem_desc_mn(sp + LY::B0) + koff;
                    const uint32_t accum = (i | k) != 0 ? 1u : 0u;
                    tc_mma(tmem_base, da0, db0, id256, accum);
                    if (MODE == MODE_WFG) tc_mma(tmem_base + 256, da1, db0, id256, accum);
                    else tc_mma(tmem_base + 256, da1, smem_desc_mn(sp + LY::B1) + koff, id128, accum);
                }
                tc_commit(&empty_bar[st]);
            }
            tc_commit(&acc_full);
        }
    } else if (my_blocks > 0) {
        // epilogue: thread = (output row m = TMEM lane, 64-column quarter); partial sums -> fp32 gradients with 16-byte atomics
        const int q = warp & 3, qd = (warp - 2) >> 2;
        const int m = q * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        mbar_wait(&acc_full, 0, a.err);
        tc_fence_after();
#pragma unroll 1
        for (int acc = 0; acc < 2; ++acc) {
            const int ncols = (MODE == MODE_WLD && acc == 1) ? 128 : 256;
            float *dst = (acc == 0 ? a.dW0 : a.dW1) + (size_t)m * ncols;
            const bool row_ok = acc == 0 ? m < a.rows0 : true;
            const int per = ncols / 4;                            // columns per quarter
#pragma unroll 1
            for (int c0 = qd * per; c0 < (qd + 1) * per; c0 += 32) {
                float v[32];
                tc_ld32(tlane + acc * 256 + c0, v);
                tc_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) atomicAdd(reinterpret_cast<float4 *>(dst + c0 + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Per-sentence column sums of a bf16 (M, cols) matrix over the rows tau >= tau_min of each sentence, atomically added into
// out (n, cols) (or out (cols) when per_sentence == 0): bias and global-condition gradients of the fused backward path.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16 *__restrict__ in, float *__restrict__ out, int T0, int cols, int tau_min, int per_sentence, int CH) {
    // thread -> (row lane, 8 columns); 4 independent 16-byte loads in flight per thread; one shared-memory transpose-reduce per block
    const int tpr = cols >> 3, rpp = blockDim.x / tpr, lr = threadIdx.x / tpr, c = (threadIdx.x - lr * tpr) * 8;
    const int n = blockIdx.y, t0 = max((int)blockIdx.x * CH, tau_min), t1 = min(T0, (int)(blockIdx.x + 1) * CH);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bf16 *base = in + (size_t)n * T0 * cols + c;
    int tau = t0 + lr;
    for (; tau + 3 * rpp < t1; tau += 4 * rpp) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4 *>(base + (size_t)(tau + u * rpp) * cols);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&w[i]);
                s[2 * i] += __low2float(b);
                s[2 * i + 1] += __high2float(b);
            }
        }
    }
    for (; tau < t1; tau += rpp) {
        const uint4 v = *reinterpret_cast<const uint4 *>(base + (size_t)tau * cols);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&w[i]);
            s[2 * i] += __low2float(b);
            s[2 * i + 1] += __high2float(b);
        }
    }
    __shared__ float red[8][257];
#pragma unroll
    for (int i = 0; i < 8; ++i) red[i][threadIdx.x] = s[i];
    __syncthreads();
    // cols outputs, each the sum over the rpp row lanes: thread j < cols handles column j
    if ((int)threadIdx.x < cols) {
        const int grp = threadIdx.x >> 3, i = threadIdx.x & 7;
        float tot = 0.f;
        for (int rr = 0; rr < rpp; ++rr) tot += red[i][rr * tpr + grp];
        atomicAdd(out + (per_sentence ? (size_t)n * cols : 0) + threadIdx.x, tot);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Causal layer (scalar input) as tensor-core GEMMs.  The 32-tap FIR x0[row][r] = sum_k wav[n, tau+k] * Wc[k][r] is a GEMM over
// the im2col matrix of the waveform; to keep 16-bit audio exact in bf16 operands every sample is split into hi + lo bf16 parts:
//   Wcol (M, 2*ifw) = [ hi(wav[tau+0..ifw)) | lo(wav[tau+0..ifw)) ],   WcDup (2*ifw, R) = [Wc ; Wc]   ->   x0 = Wcol . WcDup
// and the weight gradient is dWc[k] = (Wcol^T . dx0)[k] + (Wcol^T . dx0)[ifw + k].
__global__ void wav_im2col_kernel(const float *__restrict__ wav, bf16 *__restrict__ Wcol, int N, int Tlen, int T0, int ifw) {
    const long total = (long)N * T0 * ifw;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int k = (int)(e % ifw);
        const long row = e / ifw;
        const int n = (int)(row / T0), tau = (int)(row - (long)n * T0);
        const float x = wav[(size_t)n * Tlen + tau + k];
        const bf16 hi = __float2bfloat16_rn(x);
        Wcol[row * (2 * ifw) + k] = hi;
        Wcol[row * (2 * ifw) + ifw + k] = __float2bfloat16_rn(x - __bfloat162float(hi));
    }
}
__global__ void causal_dup_kernel(const float *__restrict__ Wc, bf16 *__restrict__ WcDup, int ifw, int R) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ifw * R) {
        const bf16 v = __float2bfloat16_rn(Wc[i]);
        WcDup[i] = v;
        WcDup[ifw * R + i] = v;
    }
}
__global__ void causal_fold_kernel(const float *__restrict__ tmp, float *__restrict__ dWc, int ifw, int R) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ifw * R) dWc[i] = tmp[i] + tmp[ifw * R + i];
}

// K-major (transposed) bf16 copies of the per-layer kernels, rebuilt whenever the parameters change:
//   WfgT (L*256, 384): row = l*256 + n (n: filter 0..127 | gate 128..255), column k: [0,128) tap x[tau-d], [128,256) tap x[tau],
//                      [256, 256+C) local condition, zero beyond;   WdT (L*128, 128): row = l*128 + r, column = d.
//   WdP (L*128, 128): Wd as stored (row = l*128 + d, column = r): B operand of dz = dx_next . Wd^T;
//   WdxP (L*128, 512): row = l*128 + r, columns [0,256) = Wfg[1][r][:], [256,512) = Wfg[0][r][:]: B operand of the dx GEMM.
__global__ void transpose_weights_kernel(const float *__restrict__ P, long o_layer_w, long stride, long o_wfg, long o_wlc, long o_wd, int L,
                                         int C, bf16 *__restrict__ WfgT, bf16 *__restrict__ WdT, bf16 *__restrict__ WdP, bf16 *__restrict__ WdxP) {
    const long tot1 = (long)L * NFG * KTOT, tot2 = (long)L * ND * ND, tot3 = (long)L * ND * 512;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot1 + 2 * tot2 + tot3; e += (long)gridDim.x * blockDim.x) {
        if (e >= tot1 + tot2) {
            if (e < tot1 + 2 * tot2) {
                const long e3 = e - tot1 - tot2;
                const long l = e3 / (ND * ND);
                WdP[e3] = __float2bfloat16_rn(P[o_layer_w + stride * l + o_wd + (e3 - l * ND * ND)]);
            } else {
                const long e4 = e - tot1 - 2 * tot2;
                const int k = (int)(e4 % 512);
                const long rr = e4 / 512;
                const int r = (int)(rr % ND);
                const long l = rr / ND;
                const int tap = k < 256 ? 1 : 0, cc = k & 255;
                WdxP[e4] = __float2bfloat16_rn(P[o_layer_w + stride * l + o_wfg + ((long)tap * ND + r) * NFG + cc]);
            }
            continue;
        }
        if (e < tot1) {
            const int k = (int)(e % KTOT);
            const long rn = e / KTOT;
            const int n = (int)(rn % NFG), l = (int)(rn / NFG);
            const float *W = P + o_layer_w + stride * l;
            float v = 0.f;
            if (k < 256) v = W[o_wfg + (long)k * NFG + n];               // (2, R, 2D) flattened: row k = tap*128 + r
            else if (k - 256 < C) v = W[o_wlc + (long)(k - 256) * NFG + n];
            WfgT[e] = __float2bfloat16_rn(v);
        } else {
            const long e2 = e - tot1;
            const int dch = (int)(e2 % ND);
            const long rr = e2 / ND;
            const int r = (int)(rr % ND), l = (int)(rr / ND);
            WdT[e2] = __float2bfloat16_rn(P[o_layer_w + stride * l + o_wd + (long)dch * ND + r]);   // Wd is (D, R) row-major
        }
    }
}

}  // namespace wntf
