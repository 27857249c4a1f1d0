// taco_gemm_tc.cuh -- tf.layers.conv1d('same') / dense of the CBHG stacks (modules.py:25-57, 91-96) on the 5th-gen tensor
// cores: tcgen05.mma kind::tf32 with a three-product split that keeps fp32 accuracy.
//
//   x = x_hi + x_lo,  x_hi = rna_tf32(x),  x_lo = rna_tf32(x - x_hi)        (|x_lo| <= 2^-11 |x|, both exactly representable)
//   A.W  ~=  A_hi.W_hi + A_hi.W_lo + A_lo.W_hi                              (dropped: A_lo.W_lo <= 2^-22 |A||W|), fp32 accumulation in TMEM
//
// which is the accuracy of an fp32 FMA chain to within a few ulp -- north_star's 1e-4 on the mel / linear outputs holds with
// two orders of magnitude to spare (tests/test_taco_gpu.py compares against the fp64-accumulating oracle).
//
// Implicit GEMM: one CTA = 128 time steps of one sentence x NT output channels of one problem.  For tap j and channel block c0
// the A tile is the TMA box (32 channels, 128 steps, 1 sentence) at (c0, t0 + j - pl, b) of the (Cip, T, B) view of the split
// input: steps outside [0, T) are zero-filled by the TMA unit, which IS the 'same' padding.  The B tile is the box (32, NT) of
// the transposed, split, zero-padded weights (N, ktaps*Cip).  Both land 128-byte swizzled; 4 k-steps of 8 per tile, 3 MMAs each.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue (TMEM -> registers
// -> bias / activation / batch-norm / residual / per-sentence row vector -> global), the same epilogue order as taco_gemm_kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace taco {
namespace tc {

constexpr int TM = 128;              // time steps per CTA
constexpr int TK = 32;               // channels per k-tile (128 bytes: one swizzle atom)
constexpr int THREADS = 192;
constexpr int MAX_PROBS = 16;

struct Prob {                        // one conv / dense of a launch (blockIdx.z)
    const float *bias, *bn_scale, *bn_shift, *R, *rowvec;
    float *C;
    int ldc, ldr, ldrv;
    int ktaps, pl, N, act, epi;      // epi 1: highway, columns (2c, 2c+1) = (H_c, T_c), R = the layer input (modules.py:83-89)
    int b_row0;                      // first row of this problem's weights in the B tensor
};
struct Args {
    Prob p[MAX_PROBS];
    int T, B, Cip;
    unsigned *err;
    long long *dbg;                  // TACO_TC_DEBUG: 16 clock64 stamps of CTA (0,0,0) of this launch, or null
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a protocol bug must fail the launch, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, unsigned *err) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) {
            if (err) atomicExch(err, 1u);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, fp32 containers read as tf32, both operands K-major
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
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

// K-major operand, 128-byte swizzle, one swizzle atom (32 fp32) along K: start address >> 4 in bits [0,14), stride byte offset
// (8 rows x 128 B) >> 4 in [32,46), descriptor version 1 in [46,48), SWIZZLE_128B (2) in [61,64)  (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major (bits 15, 16 = 0),
// N >> 3 in [17,23), M >> 4 in [24,29)
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---- operand preparation ------------------------------------------------------------------------------------------------
// Activations: X (B*T rows, lda) -> hi / lo (B*T, Cip), channels >= Ci zero.  pool: max_pooling1d(2, 1, 'same') folded in
// (max(X[t], X[t+1]), X[t] at the last step of the sentence; modules.py:40).
__global__ void split_act_kernel(const float *__restrict__ X, int lda, int Ci, int Cip, int T, long rows, int pool, float *__restrict__ hi,
                                 float *__restrict__ lo) {
    const long total = rows * (Cip / 4);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / (Cip / 4);
        const int c = (int)(i - r * (Cip / 4)) * 4;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float x = 0.0f;
            if (c + k < Ci) {
                x = __ldg(X + r * lda + c + k);
                if (pool && (r % T) + 1 < T) x = fmaxf(x, __ldg(X + (r + 1) * lda + c + k));
            }
            v[k] = x;
        }
        float4 h, l;
        h.x = rna_tf32(v[0]); h.y = rna_tf32(v[1]); h.z = rna_tf32(v[2]); h.w = rna_tf32(v[3]);
        l.x = rna_tf32(v[0] - h.x); l.y = rna_tf32(v[1] - h.y); l.z = rna_tf32(v[2] - h.z); l.w = rna_tf32(v[3] - h.w);
        *reinterpret_cast<float4 *>(hi + r * Cip + c) = h;
        *reinterpret_cast<float4 *>(lo + r * Cip + c) = l;
    }
}
// Weights: W (ktaps*Ci, N) row-major (TF (k, Ci, Co)) -> hi / lo (N rows starting at row0, Kp_stride), element (n, j*Cip + ci)
__global__ void split_weight_kernel(const float *__restrict__ W, int ktaps, int Ci, int Cip, int N, int row0, int Kp_stride, float *__restrict__ hi,
                                    float *__restrict__ lo) {
    const long total = (long)N * ktaps * Ci;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const long k = i / N;
        const int j = (int)(k / Ci), ci = (int)(k - (long)j * Ci);
        const float x = __ldg(W + i);
        const float h = rna_tf32(x);
        const size_t o = (size_t)(row0 + n) * Kp_stride + (size_t)j * Cip + ci;
        hi[o] = h;
        lo[o] = rna_tf32(x - h);
    }
}

// ---- the GEMM ------------------------------------------------------------------------------------------------------------
template <int NT>
struct Layout {
    static constexpr int STAGES = NT > 128 ? 2 : 3;
    static constexpr int A_BYTES = TM * TK * 4, B_BYTES = NT * TK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
    static constexpr int TMEM_COLS = NT <= 32 ? 32 : (NT <= 64 ? 64 : (NT <= 128 ? 128 : 256));
};

template <int ACT>
__device__ __forceinline__ float act_apply(float v) {
    if (ACT == 1) return fmaxf(v, 0.0f);
    if (ACT == 2) return 1.0f / (1.0f + expf(-v));
    if (ACT == 3) return tanhf(v);
    if (ACT == 4) return v / (fabsf(v) + 1.0f);
    return v;
}
// rows [0, rows_here) of one transposed 32-column chunk, lane = column: bias -> activation -> batch norm -> + residual -> + row vector.
// The activation is a template parameter: with a runtime switch ptxas if-converts the body and evaluates exp / tanh / two IEEE
// divisions for every element (measured: 9 k cycles per chunk instead of 1.5 k).
template <int ACT>
__device__ __forceinline__ void epi_rows(const float *tile, int lane, int rows_here, float bv, bool bn, float sc, float sh, float rv,
                                         const float *__restrict__ Rrow, int ldr, float *__restrict__ Crow, int ldc) {
#pragma unroll 1
    for (int r0 = 0; r0 < rows_here; r0 += 8) {
        float res[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) res[u] = (Rrow && r0 + u < rows_here) ? __ldg(Rrow + (size_t)(r0 + u) * ldr) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int rr = r0 + u;
            if (rr < rows_here) {
                float x = act_apply<ACT>(tile[rr * 33 + lane] + bv);
                if (bn) x = x * sc + sh;
                Crow[(size_t)rr * ldc] = (x + res[u]) + rv;
            }
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo, const __grid_constant__ CUtensorMap map_bhi,
               const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ Args args) {
    using L = Layout<NT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // pointer arithmetic keeps the shared address space (LDS / STS)
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::STAGES * L::STAGE_BYTES);
    uint64_t *empty = full + L::STAGES;
    uint64_t *accum = empty + L::STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tdbg0 = clock64();
    long long *dbg = (args.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? args.dbg : nullptr;
    const Prob &P = args.p[blockIdx.z];
    const int tblocks = (args.T + TM - 1) / TM;
    const int b = blockIdx.x / tblocks, t0 = (blockIdx.x - b * tblocks) * TM;
    const int n0 = blockIdx.y * NT;
    if (n0 >= P.N) return;
    const int cblocks = args.Cip / TK;
    const int nkt = P.ktaps * cblocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < L::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(L::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (dbg && threadIdx.x == 0) { dbg[0] = nkt; dbg[1] = clock64() - tdbg0; }

    if (warp == 0) {
        if (lane == 0) {
            for (int kt = 0; kt < nkt; ++kt) {
                const int s = kt % L::STAGES, it = kt / L::STAGES;
                if (it > 0) mbar_wait(&empty[s], (uint32_t)(it - 1) & 1u, args.err);
                uint8_t *st = smem + s * L::STAGE_BYTES;
                const int j = kt / cblocks, c0 = (kt - j * cblocks) * TK;
                mbar_expect_tx(&full[s], (uint32_t)L::STAGE_BYTES);
                tma_load_3d(st, &map_ahi, &full[s], c0, t0 + j - P.pl, b);
                tma_load_3d(st + L::A_BYTES, &map_alo, &full[s], c0, t0 + j - P.pl, b);
                tma_load_2d(st + 2 * L::A_BYTES, &map_bhi, &full[s], kt * TK, P.b_row0 + n0);
                tma_load_2d(st + 2 * L::A_BYTES + L::B_BYTES, &map_blo, &full[s], kt * TK, P.b_row0 + n0);
                if (dbg && kt < 2) dbg[2 + kt] = clock64() - tdbg0;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_tf32(TM, NT);
            for (int kt = 0; kt < nkt; ++kt) {
                const int s = kt % L::STAGES, it = kt / L::STAGES;
                mbar_wait(&full[s], (uint32_t)it & 1u, args.err);
                if (dbg && (kt < 2 || kt == nkt - 1)) dbg[kt < 2 ? 4 + kt : 6] = clock64() - tdbg0;
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
                const uint64_t d_ahi = smem_desc(sa), d_alo = smem_desc(sa + L::A_BYTES), d_bhi = smem_desc(sa + 2 * L::A_BYTES),
                               d_blo = smem_desc(sa + 2 * L::A_BYTES + L::B_BYTES);
#pragma unroll
                for (int ks = 0; ks < TK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 32 >> 4);          // 8 fp32 = 32 bytes along K inside the swizzle atom
                    tc_mma_tf32(tmem, d_alo + adv, d_bhi + adv, idesc, (kt | ks) ? 1u : 0u);
                    tc_mma_tf32(tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
                    tc_mma_tf32(tmem, d_ahi + adv, d_bhi + adv, idesc, 1u);
                }
                tc_commit(&empty[s]);                                       // frees the stage when these MMAs have read it
            }
            tc_commit(accum);
        }
    } else {
        // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 = time steps t0 + that range.  Per 32-column chunk: lane = row does the
        // per-column part (bias, activation, batch norm; highway gate) on its TMEM row, the chunk is transposed through a padded
        // shared tile (the pipeline stages are free by now), and lane = column adds the residual / per-sentence vector and stores:
        // every global access of the warp is one contiguous 128-byte row segment.
        const int part = warp & 3;
        mbar_wait(accum, 0u, args.err);
        if (dbg && warp == 2 && lane == 0) dbg[7] = clock64() - tdbg0;
        tc_fence_after();
        float *tile = reinterpret_cast<float *>(smem) + part * (32 * 33);
        const int tw0 = t0 + part * 32;
        const size_t m0 = (size_t)b * args.T + tw0;
        const bool hw = P.epi == 1;
        const float *__restrict__ Rp = P.R;
        const float *__restrict__ rowv = P.rowvec ? P.rowvec + (size_t)b * P.ldrv : nullptr;
        float *__restrict__ Cp = P.C;
        const int ldr = P.ldr, ldc = P.ldc, act = P.act, Nn = P.N;
        const int rows_here = min(32, args.T - tw0);                        // <= 0: this warp's rows are all past the sentence
#pragma unroll 1
        for (int c = 0; c < NT; c += 32) {
            if (n0 + c >= Nn) break;                                        // warp-uniform
            {
                float v[32];
                tc_ld32(tmem + ((uint32_t)(part * 32) << 16) + (uint32_t)c, v);
                tc_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) tile[lane * 33 + e] = v[e];
            }
            __syncwarp();
            if (hw) {
                // columns (2l, 2l+1) of the chunk = (H, T) of output channel (n0 + c)/2 + l: lanes 0-15 take even rows, 16-31 odd rows
                const int l = lane & 15, oc = ((n0 + c) >> 1) + l, n = n0 + c + 2 * l;
                if (n + 1 < Nn) {
                    const float bh = __ldg(P.bias + n), bt = __ldg(P.bias + n + 1);
#pragma unroll 1
                    for (int r0 = lane >> 4; r0 < rows_here; r0 += 8) {
                        float xin[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) xin[u] = (r0 + 2 * u < rows_here) ? __ldg(Rp + (m0 + r0 + 2 * u) * ldr + oc) : 0.0f;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int rr = r0 + 2 * u;
                            if (rr < rows_here) {
                                const float hh = fmaxf(tile[rr * 33 + 2 * l] + bh, 0.0f);
                                const float tg = 1.0f / (1.0f + expf(-(tile[rr * 33 + 2 * l + 1] + bt)));
                                Cp[(m0 + rr) * ldc + oc] = hh * tg + xin[u] * (1.0f - tg);
                            }
                        }
                    }
                }
            } else {
                const int nn = n0 + c + lane;
                if (nn < Nn) {
                    const float bv = P.bias ? __ldg(P.bias + nn) : 0.0f;
                    const float sc = P.bn_scale ? __ldg(P.bn_scale + nn) : 1.0f, sh = P.bn_scale ? __ldg(P.bn_shift + nn) : 0.0f;
                    const float rv = rowv ? __ldg(rowv + nn) : 0.0f;
                    const bool bn = P.bn_scale != nullptr;
                    const float *Rrow = Rp ? Rp + m0 * ldr + nn : nullptr;
                    float *Crow = Cp + m0 * ldc + nn;
                    switch (act) {                                          // warp-uniform
                        case 1: epi_rows<1>(tile, lane, rows_here, bv, bn, sc, sh, rv, Rrow, ldr, Crow, ldc); break;
                        case 2: epi_rows<2>(tile, lane, rows_here, bv, bn, sc, sh, rv, Rrow, ldr, Crow, ldc); break;
                        case 3: epi_rows<3>(tile, lane, rows_here, bv, bn, sc, sh, rv, Rrow, ldr, Crow, ldc); break;
                        case 4: epi_rows<4>(tile, lane, rows_here, bv, bn, sc, sh, rv, Rrow, ldr, Crow, ldc); break;
                        default: epi_rows<0>(tile, lane, rows_here, bv, bn, sc, sh, rv, Rrow, ldr, Crow, ldc); break;
                    }
                }
            }
            __syncwarp();
            if (dbg && warp == 2 && lane == 0 && c < 64) dbg[8 + c / 32] = clock64() - tdbg0;
        }
        tc_fence_before();
    }
    __syncthreads();
    if (dbg && threadIdx.x == 64) dbg[10] = clock64() - tdbg0;
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(L::TMEM_COLS) : "memory");
    }
}

}  // namespace tc
}  // namespace taco
