// taco_kernels.cuh -- sm_100a kernels of the Tacotron text->mel path.
//
//  taco_gemm_kernel     fp32 implicit-GEMM for tf.layers.dense / tf.layers.conv1d('same') with fused epilogues
//                       (bias, activation, inference batch-norm, residual, per-sentence row vector, highway gate,
//                       max_pooling1d(2,1,'same') folded into the A loader).  modules.py:15-23,25-57,83-96.
//  taco_bigru_kernel    recurrent half of tf.nn.bidirectional_dynamic_rnn(GRUCell) (modules.py:66-73): one CTA per
//                       (sentence, direction), recurrent weights resident in shared memory.
//  taco_decoder_kernel  tf.contrib.seq2seq.dynamic_decode over the decoder cell stack (tacotron.py:151-201) as ONE
//                       persistent cooperative kernel: weight-stationary column slices in shared memory,
//                       feature-major activations, grid-wide barrier between dependent phases.
//
// Arithmetic is plain fp32 FMA with fixed summation orders and CUDA's accurate expf/tanhf/logf (no fast-math).  The large
// conv / dense contractions of the two CBHG stacks run on the tensor cores instead (taco_gemm_tc.cuh: tcgen05 kind::tf32 with a
// three-product split that keeps fp32 accuracy); taco_gemm_kernel below serves the small problems and TACO_NO_TC=1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace taco {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_TANH = 3, ACT_SOFTSIGN = 4 };
enum { EPI_LINEAR = 0, EPI_HIGHWAY = 1 };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.0f);
        case ACT_SIGMOID: return sigmoidf_(v);
        case ACT_TANH: return tanhf(v);
        case ACT_SOFTSIGN: return v / (fabsf(v) + 1.0f);
        default: return v;
    }
}

// ---------------------------------------------------------------------------------------------------
// Implicit GEMM.  C[m, n] = epi( sum_{j<ktaps} sum_{ci<Ci} X[b, t - pl + j, ci] * W[j*Ci + ci, n] ), m = b*T + t.
struct GemmProb {
    const float *A;          // (B*T, lda) rows, already offset to the first input column
    const float *W;          // (ktaps*Ci, N) row-major == TF (k, Ci, Co) / (Ci, Co)
    const float *bias;       // (N) or null
    const float *bn_scale;   // (N) or null: v*scale + shift AFTER the activation (modules.py:95-96)
    const float *bn_shift;
    const float *R;          // residual (B*T, ldr) or null; for EPI_HIGHWAY the layer input
    const float *rowvec;     // (B, ldrv) added to every time step of sentence b, or null (modules.py:47-51)
    float *C;                // (B*T, ldc) rows, already offset to the first output column
    int lda, ldr, ldrv, ldc;
    int Ci, ktaps, pl, pool; // pool: read max(X[t], X[t+1]) (X[t] at the last step) instead of X[t]
    int N, act, epi;
};

constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;
constexpr int GAS = GBM + 4;     // As row stride (floats); keeps 16 B alignment of float4 reads

__global__ void __launch_bounds__(GTHREADS) taco_gemm_kernel(const GemmProb *__restrict__ probs, int B, int T) {
    const GemmProb p = probs[blockIdx.z];
    const int n0 = blockIdx.x * GBN;
    if (n0 >= p.N) return;
    const int M = B * T;
    const int m0 = blockIdx.y * GBM;
    const int K = p.ktaps * p.Ci;
    const int nkt = (K + GBK - 1) / GBK;

    __shared__ __align__(16) float As[2][GBK][GAS];
    __shared__ __align__(16) float Bs[2][GBK][GBN];

    const int tid = threadIdx.x;
    const int a_k = tid & 15, a_m = tid >> 4;           // A loader: k within tile, first of 8 rows (stride 16)
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;      // B loader: one float4
    const int tx = tid & 15, ty = tid >> 4;              // compute: cols tx*4.., rows ty*8..

    int tt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + a_m + 16 * i;
        tt[i] = (m < M) ? (m % T) : -(1 << 28);
    }
    const bool w_vec = ((p.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0);

    float ra[8];
    float4 rb;
    auto load_tile = [&](int kt) {
        const int kk = kt * GBK + a_k;
        const bool vk = kk < K;
        int j = 0, ci = 0;
        if (vk) { j = kk / p.Ci; ci = kk - j * p.Ci; }
        const int dj = j - p.pl;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int tau = tt[i] + dj;
            float v = 0.0f;
            if (vk && tau >= 0 && tau < T) {
                const float *src = p.A + (size_t)(m0 + a_m + 16 * i + dj) * p.lda + ci;
                v = __ldg(src);
                if (p.pool && tau + 1 < T) v = fmaxf(v, __ldg(src + p.lda));
            }
            ra[i] = v;
        }
        const int kb = kt * GBK + b_k;
        const int n = n0 + b_n;
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kb < K) {
            const float *src = p.W + (size_t)kb * p.N + n;
            if (w_vec && n + 3 < p.N) {
                rb = __ldg(reinterpret_cast<const float4 *>(src));
            } else {
                if (n + 0 < p.N) rb.x = __ldg(src + 0);
                if (n + 1 < p.N) rb.y = __ldg(src + 1);
                if (n + 2 < p.N) rb.z = __ldg(src + 2);
                if (n + 3 < p.N) rb.w = __ldg(src + 3);
            }
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][a_k][a_m + 16 * i] = ra[i];
        *reinterpret_cast<float4 *>(&Bs[buf][b_k][b_n]) = rb;
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) load_tile(kt + 1);
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (kt + 1 < nkt) store_tile(buf ^ 1);
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
        const int b = m / T;
        if (p.epi == EPI_HIGHWAY) {
            // columns (2c, 2c+1) = (H_c, T_c): out = relu(H)*sigmoid(T) + x*(1 - sigmoid(T))   (modules.py:83-89)
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                const int n = n0 + tx * 4 + 2 * jp;
                if (n + 1 < p.N) {
                    const float hh = fmaxf(acc[i][2 * jp] + p.bias[n], 0.0f);
                    const float tg = sigmoidf_(acc[i][2 * jp + 1] + p.bias[n + 1]);
                    const int c = n >> 1;
                    const float x = p.R[(size_t)m * p.ldr + c];
                    p.C[(size_t)m * p.ldc + c] = hh * tg + x * (1.0f - tg);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n < p.N) {
                    float v = acc[i][j];
                    if (p.bias) v += p.bias[n];
                    v = apply_act(v, p.act);
                    if (p.bn_scale) v = v * p.bn_scale[n] + p.bn_shift[n];
                    if (p.R) v += p.R[(size_t)m * p.ldr + n];
                    if (p.rowvec) v += p.rowvec[(size_t)b * p.ldrv + n];
                    p.C[(size_t)m * p.ldc + n] = v;
                }
            }
        }
    }
}

// tf.nn.embedding_lookup with row 0 forced to zeros when zero_row0 (tacotron.py:51-58).
__global__ void taco_embed_kernel(const int32_t *__restrict__ ids, const float *__restrict__ table, int n_rows, int width,
                                  int table_rows, int zero_row0, float *__restrict__ out) {
    const size_t total = (size_t)n_rows * width;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / width), c = (int)(i - (size_t)r * width);
        int id = ids[r];
        id = id < 0 ? 0 : (id >= table_rows ? table_rows - 1 : id);
        out[i] = (zero_row0 && id == 0) ? 0.0f : table[(size_t)id * width + c];
    }
}

// ---------------------------------------------------------------------------------------------------
// Bidirectional GRU, recurrent half.  XP (N, T, 2, 3U) holds x*W_x + b for [gates(2U) | candidate(U)] of each
// direction (one GEMM).  Per step: g = sigmoid(XPg + h*Wgh); r,u = split(g); c = tanh(XPc + (r*h)*Wch);
// h' = u*h + (1-u)*c.  Past `len` the output is zero and the state is carried (dynamic_rnn sequence_length);
// the backward direction starts at len-1.
struct RnnParams {
    const float *XP;           // (N, T, 2, 3U)
    const float *Wgh[2];       // (U, 2U) per direction
    const float *Wch[2];       // (U, U)
    const float *init;         // (N, 2U) [fw | bw] or null
    const int32_t *lengths;    // device (N) or null (= T)
    float *out;                // (N, T, 2U)
    int N, T, U, w_in_smem;
};

__global__ void __launch_bounds__(256) taco_bigru_kernel(const RnnParams p) {
    extern __shared__ __align__(16) float rsm[];
    const int U = p.U, U2 = 2 * p.U;
    const int dir = blockIdx.x & 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    float *h_s = rsm;              // U
    float *rh_s = h_s + U;         // U
    float *u_s = rh_s + U;         // U
    float *w_s = u_s + U;          // optional U*2U + U*U
    const float *Wg = p.Wgh[dir], *Wc = p.Wch[dir];
    if (p.w_in_smem) {
        for (int i = tid; i < U * U2; i += nt) w_s[i] = __ldg(Wg + i);
        for (int i = tid; i < U * U; i += nt) w_s[U * U2 + i] = __ldg(Wc + i);
        Wg = w_s;
        Wc = w_s + U * U2;
    }
    for (int row = blockIdx.x >> 1; row < p.N; row += gridDim.x >> 1) {
        const int len = p.lengths ? min(max(p.lengths[row], 0), p.T) : p.T;
        __syncthreads();
        for (int i = tid; i < U; i += nt) h_s[i] = p.init ? p.init[(size_t)row * U2 + dir * U + i] : 0.0f;
        // zero tail of the output
        for (int i = tid; i < (p.T - len) * U; i += nt) {
            const int t = len + i / U, c = i - (i / U) * U;
            p.out[((size_t)row * p.T + t) * U2 + dir * U + c] = 0.0f;
        }
        __syncthreads();
        for (int s = 0; s < len; ++s) {
            const int t = dir ? (len - 1 - s) : s;
            const float *xp = p.XP + (((size_t)row * p.T + t) * 2 + dir) * 3 * U;
            for (int c = tid; c < U2; c += nt) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                int k = 0;
                for (; k + 3 < U; k += 4) {
                    const float4 hv = *reinterpret_cast<const float4 *>(&h_s[k]);
                    a0 = fmaf(hv.x, Wg[(k + 0) * U2 + c], a0);
                    a1 = fmaf(hv.y, Wg[(k + 1) * U2 + c], a1);
                    a2 = fmaf(hv.z, Wg[(k + 2) * U2 + c], a2);
                    a3 = fmaf(hv.w, Wg[(k + 3) * U2 + c], a3);
                }
                for (; k < U; ++k) a0 = fmaf(h_s[k], Wg[k * U2 + c], a0);
                const float g = sigmoidf_(xp[c] + ((a0 + a1) + (a2 + a3)));
                if (c < U) rh_s[c] = g * h_s[c];
                else u_s[c - U] = g;
            }
            __syncthreads();
            float hn_keep = 0.f;
            for (int c = tid; c < U; c += nt) {        // U <= blockDim: at most one column per thread
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                int k = 0;
                for (; k + 3 < U; k += 4) {
                    const float4 hv = *reinterpret_cast<const float4 *>(&rh_s[k]);
                    a0 = fmaf(hv.x, Wc[(k + 0) * U + c], a0);
                    a1 = fmaf(hv.y, Wc[(k + 1) * U + c], a1);
                    a2 = fmaf(hv.z, Wc[(k + 2) * U + c], a2);
                    a3 = fmaf(hv.w, Wc[(k + 3) * U + c], a3);
                }
                for (; k < U; ++k) a0 = fmaf(rh_s[k], Wc[k * U + c], a0);
                const float cand = tanhf(xp[U2 + c] + ((a0 + a1) + (a2 + a3)));
                const float u = u_s[c];
                hn_keep = u * h_s[c] + (1.0f - u) * cand;
                p.out[((size_t)row * p.T + t) * U2 + dir * U + c] = hn_keep;
            }
            __syncthreads();
            if (tid < U) h_s[tid] = hn_keep;
            __syncthreads();
        }
    }
}

// Same recurrence with the recurrent weights in REGISTERS (U = 128: thread c of 256 owns gate column c, 128
// registers, and half a candidate column, 64 registers), the next step's input projections prefetched during the
// current step and a double-buffered state: three block barriers and no shared-memory weight traffic per step.
template <int U>
__global__ void __launch_bounds__(2 * U, 1) taco_bigru_reg_kernel(const RnnParams p) {
    constexpr int U2 = 2 * U, UH = U / 2;
    __shared__ __align__(16) float h_s[2][U];
    __shared__ __align__(16) float rh_s[U];
    __shared__ float u_s[U];
    __shared__ float cpart[U];
    const int dir = blockIdx.x & 1;
    const int tid = threadIdx.x;
    const int cc = tid & (U - 1), half = tid / U;
    float wg[U], wc[UH];
    {
        const float *Wg = p.Wgh[dir], *Wc = p.Wch[dir];
#pragma unroll
        for (int k = 0; k < U; ++k) wg[k] = __ldg(Wg + (size_t)k * U2 + tid);
#pragma unroll
        for (int k = 0; k < UH; ++k) wc[k] = __ldg(Wc + (size_t)(half * UH + k) * U + cc);
    }
    for (int row = blockIdx.x >> 1; row < p.N; row += gridDim.x >> 1) {
        const int len = p.lengths ? min(max(p.lengths[row], 0), p.T) : p.T;
        __syncthreads();
        if (tid < U) h_s[0][tid] = p.init ? p.init[(size_t)row * U2 + dir * U + tid] : 0.0f;
        for (int i = tid; i < (p.T - len) * U; i += U2) {
            const int t = len + i / U, c = i - (i / U) * U;
            p.out[((size_t)row * p.T + t) * U2 + dir * U + c] = 0.0f;
        }
        __syncthreads();
        float xg = 0.f, xc = 0.f;
        if (len > 0) {
            const int t0 = dir ? (len - 1) : 0;
            const float *xp = p.XP + (((size_t)row * p.T + t0) * 2 + dir) * 3 * U;
            xg = __ldg(xp + tid);
            if (half == 0) xc = __ldg(xp + U2 + cc);
        }
        int cur = 0;
        for (int s = 0; s < len; ++s) {
            const int t = dir ? (len - 1 - s) : s;
            float nxg = 0.f, nxc = 0.f;
            if (s + 1 < len) {                       // prefetch the next step's input projections
                const int tn = dir ? (t - 1) : (t + 1);
                const float *xp = p.XP + (((size_t)row * p.T + tn) * 2 + dir) * 3 * U;
                nxg = __ldg(xp + tid);
                if (half == 0) nxc = __ldg(xp + U2 + cc);
            }
            const float *hc = h_s[cur];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < U; k += 4) {
                const float4 hv = *reinterpret_cast<const float4 *>(&hc[k]);
                a0 = fmaf(hv.x, wg[k + 0], a0);
                a1 = fmaf(hv.y, wg[k + 1], a1);
                a2 = fmaf(hv.z, wg[k + 2], a2);
                a3 = fmaf(hv.w, wg[k + 3], a3);
            }
            const float g = sigmoidf_(xg + ((a0 + a1) + (a2 + a3)));
            if (half == 0) rh_s[cc] = g * hc[cc];
            else u_s[cc] = g;
            __syncthreads();
            a0 = a1 = a2 = a3 = 0.f;
#pragma unroll
            for (int k = 0; k < UH; k += 4) {
                const float4 rv = *reinterpret_cast<const float4 *>(&rh_s[half * UH + k]);
                a0 = fmaf(rv.x, wc[k + 0], a0);
                a1 = fmaf(rv.y, wc[k + 1], a1);
                a2 = fmaf(rv.z, wc[k + 2], a2);
                a3 = fmaf(rv.w, wc[k + 3], a3);
            }
            const float part = (a0 + a1) + (a2 + a3);
            if (half == 1) cpart[cc] = part;
            __syncthreads();
            if (half == 0) {
                const float cand = tanhf(xc + (part + cpart[cc]));
                const float u = u_s[cc];
                const float hn = u * hc[cc] + (1.0f - u) * cand;
                h_s[cur ^ 1][cc] = hn;
                p.out[((size_t)row * p.T + t) * U2 + dir * U + cc] = hn;
            }
            __syncthreads();
            cur ^= 1;
            xg = nxg;
            xc = nxc;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Persistent decoder.
//
// Activations live feature-major in tiles of 32 sentences: buf[tile][feature][32], so a warp whose lane is the
// sentence reads them conflict-free and every CTA stages a phase's whole input with coalesced 16 B copies.
// Every dense phase is weight-stationary: CTA c owns columns [c*ncp, (c+1)*ncp) of the phase's matrix, resident in
// shared memory for the whole launch (image built by taco_finalize).  Inside a CTA warp w stages and sums the
// K-slice [w*kper, (w+1)*kper) in increasing k; the 8 slice sums are added in increasing w, then the bias.
// Between dependent phases the grid meets at a release/acquire counter barrier; inputs that were already final
// before the previous phase (recurrent states) are copied into shared memory while the CTA waits there.
// The attention phases keep their slice of the keys / values resident in shared memory when it fits.
enum { PH_DENSE = 0, PH_ATT_SCORE = 1, PH_ATT_CTX = 2 };
enum { DE_RELU = 0, DE_LINEAR = 1, DE_GATES = 2, DE_CAND = 3, DE_OUT = 4, DE_QUERY = 5 };
enum { DB_X = 0, DB_P0, DB_P1, DB_P2, DB_P3, DB_CTX, DB_HATT, DB_RH, DB_U, DB_Q, DB_O0, DB_O1, DB_O2, DB_O3, DB_O4,
       DB_H1, DB_H2, DB_H3, DB_H4, DB_COUNT };
constexpr int DEC_THREADS = 256;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int DEC_MAX_PHASES = 24;
constexpr long long DEC_WATCHDOG_CYCLES = 6000000000LL;   // ~3 s: a lost CTA aborts the launch instead of hanging the GPU

struct DecPhase {
    int kind;                 // PH_*
    int K, N;                 // dense: input features, output columns
    int ncp, pad;             // columns per CTA; row stride of the CTA's weight slice (1, 2 or a multiple of 4)
    int w_off, b_off;         // float offsets into the CTA's shared-memory image: [K][pad] then [pad]
    int nseg, seg_buf[3], seg_K[3];
    int seg_keep[3];          // 1: this slot still holds the same, unmodified buffer from the previous dense phase (tiles == 1)
    int seg_pre[3];           // 1: this slot's buffer was final before the PREVIOUS phase started: staged during the barrier
    int epi;                  // DE_*
    int out_buf;              // RELU/LINEAR: destination; GATES: unused; CAND: h buffer (in/out); OUT: DB_X
    int h_buf;                // GATES: state h (r*h -> DB_RH, u -> DB_U)
    int res_in, res_out;      // CAND: o_out = o_in + h' (ResidualWrapper), -1 = none
    int U;                    // GATES: units
    int N1;                   // columns [0, N1) take the epilogue above; [N1, N) a second one (fused phases); N1 == N: one part
    int epi2, out_buf2;       // second part: DE_RELU / DE_LINEAR into out_buf2 (N - N1 features)
};

struct DecParams {
    DecPhase ph[DEC_MAX_PHASES];
    int n_phases;
    int N, tiles, T_in, n_steps;
    int att_type, A, mem, H;  // attention units, memory width, attention-cell units
    int OD, nm;               // decoder output width (num_mels*r), num_mels
    int chunks, fslices;      // attention work split per sentence
    int img_floats;           // per-CTA image size (floats)
    int stage_floats;         // staging area (floats)
    int keys_res, vals_res;   // 1: this CTA's key chunk / value slice stays in shared memory (one item per CTA)
    const float *img;         // (grid, img_floats)
    float *buf[DB_COUNT];     // feature-major activation tiles
    float *q_row;             // (tiles*32, A) processed query, row-major for the score phase
    const float *keys;        // (N, T_in, A)
    const float *values;      // (N, T_in, mem), zero past the length
    const int32_t *lengths;   // device (N)
    const float *nv, *ab;     // (A): normed attention_v / attention_variable, attention_b / attention_bias
    float score_bias;
    const float *loc_conv_w, *loc_conv_b, *loc_w;   // (31,1,32), (32), (32,A)
    float *score;             // (N, T_in) scratch
    float *state[2];          // (N, T_in) ping-pong attention state
    const float *manual;      // (N, n_steps, T_in) or null
    float *dec_out;           // (N, n_steps, OD)
    float *align;             // (N, T_in, n_steps)
    unsigned *barrier;        // [0] arrival counter (zeroed before the launch), [1] abort flag
    long long *prof;          // optional (grid, 8) cycle counters: stage, dot, dense total, score, ctx, barrier, all
};

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");   // .cg: L2 only, never a stale L1 line
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Inclusive prefix sum of s[0..n) in place by ONE warp: each lane owns a contiguous chunk.
__device__ __forceinline__ void warp_scan_inplace(float *s, int n, int lane) {
    const int per = (n + 31) / 32;
    const int b = lane * per, e = min(n, b + per);
    float tot = 0.f;
    for (int i = b; i < e; ++i) { tot += s[i]; s[i] = tot; }
    float inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const float base = inc - tot;
    if (lane > 0) for (int i = b; i < e; ++i) s[i] += base;
}

// One warp copies rows [kb, ke) of a dense phase's input (segments are contiguous [K_s][32] blocks in global
// memory) into the staging area with 16 B cp.async; `mode` selects the slots: 0 = every slot that is not already
// valid (kept from the previous phase or prefetched), 1 = only the prefetchable slots.
__device__ __forceinline__ void stage_rows(const DecParams &P, const DecPhase &ph, float *stage, int tile, int kb, int ke, int lane,
                                           bool keep_ok, bool pre_done, int mode) {
    int koff = 0;
    for (int s = 0; s < ph.nseg; ++s) {
        const int Ks = ph.seg_K[s];
        bool want;
        if (mode == 1) want = ph.seg_pre[s] != 0;
        else want = !((ph.seg_keep[s] && keep_ok) || (ph.seg_pre[s] && pre_done));
        const int sb = max(kb, koff), se = min(ke, koff + Ks);
        if (want && sb < se) {
            const float *src = P.buf[ph.seg_buf[s]] + ((size_t)tile * Ks + (sb - koff)) * 32;
            float *dst = stage + sb * 32;
            const int n4 = (se - sb) * 8;
            for (int i = lane; i < n4; i += 32) cp_async16(dst + 4 * i, src + 4 * i);
        }
        koff += Ks;
    }
}

__global__ void __launch_bounds__(DEC_THREADS, 1) taco_decoder_kernel(const DecParams *__restrict__ Pg) {
    extern __shared__ __align__(16) float dsm[];
    __shared__ DecParams P;
    __shared__ float s_red[2];
    __shared__ int s_abort;
    {
        const int *src = reinterpret_cast<const int *>(Pg);
        int *dst = reinterpret_cast<int *>(&P);
        for (int i = threadIdx.x; i < (int)(sizeof(DecParams) / 4); i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x == 0) s_abort = 0;
    }
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;

    float *img = dsm;                                  // img_floats (rounded up to 4)
    float *red = img + ((P.img_floats + 3) & ~3);      // DEC_WARPS * 4 * 32
    float *nv_s = red + DEC_WARPS * 4 * 32;            // A
    float *ab_s = nv_s + P.A;                          // A
    float *stage = ab_s + P.A;                         // stage_floats
    float *keys_s = stage + P.stage_floats;            // chunk positions * A when keys_res
    const int cpos = (P.T_in + P.chunks - 1) / P.chunks;
    const int fs = (P.mem + P.fslices - 1) / P.fslices;
    float *vals_s = keys_s + (P.keys_res ? cpos * P.A : 0);   // T_in * fs when vals_res
    for (int i = tid; i < P.img_floats; i += DEC_THREADS) img[i] = __ldg(P.img + (size_t)cta * P.img_floats + i);
    for (int i = tid; i < P.A; i += DEC_THREADS) { nv_s[i] = __ldg(P.nv + i); ab_s[i] = __ldg(P.ab + i); }
    if (P.keys_res && cta < P.N * P.chunks) {
        const int n = cta / P.chunks, ch = cta - n * P.chunks;
        const int j0 = ch * cpos, j1 = min(P.T_in, j0 + cpos);
        const float *src = P.keys + ((size_t)n * P.T_in + j0) * P.A;
        for (int i = tid; i < (j1 - j0) * P.A; i += DEC_THREADS) keys_s[i] = __ldg(src + i);
    }
    if (P.vals_res && cta < P.N * P.fslices) {
        const int n = cta / P.fslices, sl = cta - n * P.fslices;
        const int f0 = sl * fs, nf = min(P.mem, f0 + fs) - f0;
        for (int i = tid; i < P.T_in * nf; i += DEC_THREADS) {
            const int j = i / nf, f = i - j * nf;
            vals_s[j * fs + f] = __ldg(P.values + ((size_t)n * P.T_in + j) * P.mem + f0 + f);
        }
    }
    __syncthreads();

    int staged_at = -2;    // global phase index at which this CTA last staged a dense phase's inputs
    int prefetched_for = -2;
    bool aborted = false;
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // in-kernel phase profile (cycles of thread 0), written when P.prof != null
    const long long t_begin = clock64();
    for (int step = 0; step < P.n_steps && !aborted; ++step) {
        for (int pi = 0; pi < P.n_phases; ++pi) {
            const int gpi = step * P.n_phases + pi;
            const long long t_ph = clock64();
            const DecPhase &ph = P.ph[pi];
            if (ph.kind == PH_DENSE) {
                const int c_begin = cta * ph.ncp;
                const int ncols = min(ph.ncp, ph.N - c_begin);          // may be <= 0: idle in this phase
                if (ncols > 0) {
                    const float *Ws = img + ph.w_off;
                    const float *bs = img + ph.b_off;
                    const int kper = (ph.K + DEC_WARPS - 1) / DEC_WARPS;
                    const int kb = min(ph.K, warp * kper), ke = min(ph.K, kb + kper);
                    for (int tile = 0; tile < P.tiles; ++tile) {
                        if (tile > 0) __syncthreads();            // epilogue readers of the previous tile are done
                        stage_rows(P, ph, stage, tile, kb, ke, lane, P.tiles == 1 && staged_at == gpi - 1,
                                   P.tiles == 1 && prefetched_for == gpi, 0);
                        cp_async_wait_all();
                        __syncwarp();
                        staged_at = gpi;
                        const long long t_st = clock64();
                        pt[0] += t_st - t_ph;
                        for (int c0 = 0; c0 < ncols; c0 += 4) {
                            const long long t_c0 = clock64();
                            const int cl = c0 + warp;                 // epilogue column of this warp (warps 0..3)
                            const int col = c_begin + cl;
                            const bool ev = warp < 4 && cl < ncols;
                            // operands of the epilogue that live in global memory: start the loads before the dot loop
                            float pu = 0.f, phv = 0.f, po = 0.f;
                            if (ev && ph.epi == DE_CAND) {
                                const size_t idx = ((size_t)tile * ph.N + col) * 32 + lane;
                                pu = __ldcg(&P.buf[DB_U][idx]);
                                phv = __ldcg(&P.buf[ph.out_buf][idx]);
                                if (ph.res_out >= 0) po = __ldcg(&P.buf[ph.res_in][idx]);
                            }
                            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                            if (ph.pad >= 4) {
#pragma unroll 8
                                for (int k = kb; k < ke; ++k) {
                                    const float a = stage[k * 32 + lane];
                                    const float4 w = *reinterpret_cast<const float4 *>(&Ws[k * ph.pad + c0]);
                                    a0 = fmaf(a, w.x, a0);
                                    a1 = fmaf(a, w.y, a1);
                                    a2 = fmaf(a, w.z, a2);
                                    a3 = fmaf(a, w.w, a3);
                                }
                            } else if (ph.pad == 2) {
#pragma unroll 8
                                for (int k = kb; k < ke; ++k) {
                                    const float a = stage[k * 32 + lane];
                                    const float2 w = *reinterpret_cast<const float2 *>(&Ws[k * 2]);
                                    a0 = fmaf(a, w.x, a0);
                                    a1 = fmaf(a, w.y, a1);
                                }
                            } else {
#pragma unroll 8
                                for (int k = kb; k < ke; ++k) a0 = fmaf(stage[k * 32 + lane], Ws[k], a0);
                            }
                            pt[1] += clock64() - t_c0;
                            if (c0 > 0) __syncthreads();
                            red[(warp * 4 + 0) * 32 + lane] = a0;
                            red[(warp * 4 + 1) * 32 + lane] = a1;
                            red[(warp * 4 + 2) * 32 + lane] = a2;
                            red[(warp * 4 + 3) * 32 + lane] = a3;
                            __syncthreads();
                            if (ev) {
                                float v = 0.f;
#pragma unroll
                                for (int w = 0; w < DEC_WARPS; ++w) v += red[(w * 4 + warp) * 32 + lane];
                                v += bs[cl];
                                const size_t ti = (size_t)tile * 32 + lane;   // padded sentence index
                                // fused phases (two matrices that read the same input side by side): the columns from N1 on
                                // belong to the second matrix and take its epilogue
                                int epi = ph.epi, ecol = col, eN = ph.N1, eout = ph.out_buf;
                                if (col >= ph.N1) { epi = ph.epi2; ecol = col - ph.N1; eN = ph.N - ph.N1; eout = ph.out_buf2; }
                                switch (epi) {
                                    case DE_RELU:
                                        P.buf[eout][((size_t)tile * eN + ecol) * 32 + lane] = fmaxf(v, 0.f);
                                        break;
                                    case DE_LINEAR:
                                        P.buf[eout][((size_t)tile * eN + ecol) * 32 + lane] = v;
                                        break;
                                    case DE_QUERY:
                                        P.q_row[ti * ph.N + col] = v;
                                        break;
                                    case DE_GATES: {
                                        const float g = sigmoidf_(v);
                                        if (col < ph.U) {
                                            // the state h is the last staged segment: rows [K-U, K)
                                            const float hv = stage[(ph.K - ph.U + col) * 32 + lane];
                                            P.buf[DB_RH][((size_t)tile * ph.U + col) * 32 + lane] = g * hv;
                                        } else {
                                            P.buf[DB_U][((size_t)tile * ph.U + (col - ph.U)) * 32 + lane] = g;
                                        }
                                        break;
                                    }
                                    case DE_CAND: {
                                        const size_t idx = ((size_t)tile * ph.N + col) * 32 + lane;
                                        const float c = tanhf(v);
                                        const float hn = pu * phv + (1.0f - pu) * c;
                                        P.buf[ph.out_buf][idx] = hn;
                                        if (ph.res_out >= 0) P.buf[ph.res_out][idx] = po + hn;
                                        break;
                                    }
                                    case DE_OUT: {
                                        if (ti < (size_t)P.N) P.dec_out[(ti * P.n_steps + step) * P.OD + col] = v;
                                        if (col >= P.OD - P.nm) P.buf[DB_X][((size_t)tile * P.nm + (col - (P.OD - P.nm))) * 32 + lane] = v;
                                        break;
                                    }
                                }
                            }
                        }
                    }
                }
            } else if (ph.kind == PH_ATT_SCORE) {
                // item = (sentence, chunk of positions): score_j = sum_k nv_k tanh(keys_jk + q_k [+ loc_jk] + b_k) (+ bias)
                float *q_s = stage;                 // A
                float *st_s = q_s + P.A;            // T_in (loc_sen: previous cumulative alignments)
                float *cw_s = st_s + P.T_in;        // 31*32 conv kernel + 32 bias (loc_sen)
                const int nitems = P.N * P.chunks;
                const float *st_prev = P.state[step & 1];
                for (int item = cta; item < nitems; item += G) {
                    const int n = item / P.chunks, ch = item - n * P.chunks;
                    const int len = min(max(__ldg(P.lengths + n), 0), P.T_in);
                    const int j0 = ch * cpos, j1 = min(P.T_in, j0 + cpos);
                    __syncthreads();
                    for (int k = tid; k < P.A; k += DEC_THREADS) q_s[k] = __ldcg(P.q_row + (size_t)n * P.A + k);
                    if (P.att_type == 2) {
                        for (int j = tid; j < P.T_in; j += DEC_THREADS) st_s[j] = __ldcg(st_prev + (size_t)n * P.T_in + j);
                        for (int i = tid; i < 31 * 32; i += DEC_THREADS) cw_s[i] = __ldg(P.loc_conv_w + i);
                        for (int i = tid; i < 32; i += DEC_THREADS) cw_s[31 * 32 + i] = __ldg(P.loc_conv_b + i);
                    }
                    __syncthreads();
                    for (int j = j0 + warp; j < j1; j += DEC_WARPS) {
                        float s;
                        if (j >= len) {
                            s = -INFINITY;                                  // _maybe_mask_score
                        } else {
                            float f = 0.f;
                            if (P.att_type == 2) {                          // location features: conv1d(31, 'same') of the state
                                for (int i = 0; i < 31; ++i) {
                                    const int jj = j - 15 + i;
                                    if (jj >= 0 && jj < P.T_in) f = fmaf(st_s[jj], cw_s[i * 32 + lane], f);
                                }
                                f += cw_s[31 * 32 + lane];
                            }
                            const float *kr = P.keys_res ? keys_s + (size_t)(j - j0) * P.A : P.keys + ((size_t)n * P.T_in + j) * P.A;
                            float part = 0.f;
                            for (int k = lane; k < P.A; k += 32) {
                                float e = kr[k] + q_s[k];
                                if (P.att_type == 2) {
                                    float loc = 0.f;
                                    for (int c = 0; c < 32; ++c) loc = fmaf(__shfl_sync(0xffffffffu, f, c), __ldg(P.loc_w + c * P.A + k), loc);
                                    e += loc;
                                }
                                e += ab_s[k];
                                part = fmaf(nv_s[k], tanhf(e), part);
                            }
                            s = warp_sum(part) + P.score_bias;
                        }
                        if (lane == 0) P.score[(size_t)n * P.T_in + j] = s;
                    }
                }
            } else {   // PH_ATT_CTX
                // item = (sentence, slice of context features): every item of a sentence recomputes the alignment
                // (cheap), slice 0 publishes it; context_f = sum_j a_j values_jf.
                float *sc = stage;                  // T_in: score -> p / softmax numerators -> alignment
                float *pv = sc + P.T_in;            // T_in: previous state
                float *w1 = pv + P.T_in;            // T_in: work
                float *w2 = w1 + P.T_in;            // T_in: work
                float *part = w2 + P.T_in;          // 4 * fs partial sums
                const int nitems = P.N * P.fslices;
                const float *st_prev = P.state[step & 1];
                float *st_next = P.state[(step + 1) & 1];
                for (int item = cta; item < nitems; item += G) {
                    const int n = item / P.fslices, sl = item - n * P.fslices;
                    const int T_in = P.T_in;
                    __syncthreads();
                    for (int j = tid; j < T_in; j += DEC_THREADS) {
                        sc[j] = __ldcg(P.score + (size_t)n * T_in + j);
                        pv[j] = __ldcg(st_prev + (size_t)n * T_in + j);
                    }
                    __syncthreads();
                    if (P.att_type == 2) {
                        // softmax over the masked energies; state += alignment (rnn_wrappers.py:678-689)
                        if (warp == 0) {
                            float m = -INFINITY;
                            for (int j = lane; j < T_in; j += 32) m = fmaxf(m, sc[j]);
                            m = warp_max(m);
                            if (lane == 0) s_red[0] = m;
                        }
                        __syncthreads();
                        for (int j = tid; j < T_in; j += DEC_THREADS) w1[j] = expf(sc[j] - s_red[0]);
                        __syncthreads();
                        if (warp == 0) {
                            float s = 0.f;
                            for (int j = lane; j < T_in; j += 32) s += w1[j];
                            s = warp_sum(s);
                            if (lane == 0) s_red[1] = s;
                        }
                        __syncthreads();
                        for (int j = tid; j < T_in; j += DEC_THREADS) {
                            const float a = w1[j] / s_red[1];
                            sc[j] = a;
                            if (sl == 0) st_next[(size_t)n * T_in + j] = a + pv[j];
                        }
                    } else {
                        // monotonic_attention(mode='parallel'): p = sigmoid(score);
                        // cp = exp(exclusive_cumsum(log(clip(1-p, tiny, 1)))); a = p*cp*cumsum(prev/clip(cp,1e-10,1))
                        for (int j = tid; j < T_in; j += DEC_THREADS) {
                            const float s = sc[j];
                            const float pj = (s == -INFINITY) ? 0.0f : sigmoidf_(s);
                            sc[j] = pj;
                            w1[j] = logf(fminf(fmaxf(1.0f - pj, 1.17549435e-38f), 1.0f));
                        }
                        __syncthreads();
                        if (warp == 0) warp_scan_inplace(w1, T_in, lane);
                        __syncthreads();
                        for (int j = tid; j < T_in; j += DEC_THREADS) {
                            const float cp = expf(j > 0 ? w1[j - 1] : 0.0f);
                            w2[j] = pv[j] / fminf(fmaxf(cp, 1e-10f), 1.0f);
                            pv[j] = cp;                         // previous state no longer needed
                        }
                        __syncthreads();
                        if (warp == 0) warp_scan_inplace(w2, T_in, lane);
                        __syncthreads();
                        for (int j = tid; j < T_in; j += DEC_THREADS) {
                            const float a = sc[j] * pv[j] * w2[j];
                            sc[j] = a;
                            if (sl == 0) st_next[(size_t)n * T_in + j] = a;
                        }
                    }
                    __syncthreads();
                    if (P.manual) {                              // is_manual_attention override (rnn_wrappers.py:374)
                        for (int j = tid; j < T_in; j += DEC_THREADS)
                            sc[j] = __ldg(P.manual + ((size_t)n * P.n_steps + step) * T_in + j);
                        __syncthreads();
                    }
                    if (sl == 0)
                        for (int j = tid; j < T_in; j += DEC_THREADS) P.align[((size_t)n * T_in + j) * P.n_steps + step] = sc[j];
                    // context slice
                    const int f0 = sl * fs, f1 = min(P.mem, f0 + fs);
                    const int nf = f1 - f0;
                    if (nf > 0) {
                        const int groups = max(1, min(4, DEC_THREADS / nf));
                        const int f = tid % nf, g = tid / nf;
                        const int len = min(max(__ldg(P.lengths + n), 0), T_in);
                        if (g < groups) {
                            float acc = 0.f;
                            if (P.vals_res) {
                                for (int j = g; j < len; j += groups) acc = fmaf(sc[j], vals_s[j * fs + f], acc);
                            } else {
                                const float *vp = P.values + (size_t)n * T_in * P.mem + f0 + f;
#pragma unroll 4
                                for (int j = g; j < len; j += groups) acc = fmaf(sc[j], __ldg(vp + (size_t)j * P.mem), acc);
                            }
                            part[g * nf + f] = acc;
                        }
                        __syncthreads();
                        for (int ff = tid; ff < nf; ff += DEC_THREADS) {
                            float v = part[ff];
                            for (int gg = 1; gg < groups; ++gg) v += part[gg * nf + ff];
                            P.buf[DB_CTX][((size_t)(n >> 5) * P.mem + f0 + ff) * 32 + (n & 31)] = v;
                        }
                    }
                }
            }
            // ---- grid-wide barrier: release/acquire counter; state inputs of the next phase are staged meanwhile ----
            const long long t_bar = clock64();
            if (ph.kind == PH_ATT_SCORE) pt[3] += t_bar - t_ph;
            else if (ph.kind == PH_ATT_CTX) pt[4] += t_bar - t_ph;
            else pt[2] += t_bar - t_ph;
            __syncthreads();
            if (pi + 1 < P.n_phases && P.tiles == 1) {
                const DecPhase &nx = P.ph[pi + 1];
                if (nx.kind == PH_DENSE && (nx.seg_pre[0] | nx.seg_pre[1] | nx.seg_pre[2]) && cta * nx.ncp < nx.N) {
                    const int kper = (nx.K + DEC_WARPS - 1) / DEC_WARPS;
                    const int kb = min(nx.K, warp * kper), ke = min(nx.K, kb + kper);
                    stage_rows(P, nx, stage, 0, kb, ke, lane, false, false, 1);
                    prefetched_for = gpi + 1;
                }
            }
            if (tid == 0) {
                red_release_add(P.barrier, 1u);
                const unsigned target = (unsigned)(gpi + 1) * (unsigned)G;
                const long long t0 = clock64();
                unsigned spins = 0;
                // spin on relaxed loads (~350 cycles each against ~1200 for ld.acquire.gpu, profiles/r02_probe_poll.log); one acquire
                // load of the counter after the spin orders the CTA's reads of the other CTAs' results
                while ((int)(ld_relaxed_u32(P.barrier) - target) < 0) {
                    if ((++spins & 1023u) == 0) {
                        if (ld_acquire_u32(P.barrier + 1) != 0u) { s_abort = 1; break; }
                        if (clock64() - t0 > DEC_WATCHDOG_CYCLES) { atomicExch(P.barrier + 1, 1u); s_abort = 1; break; }
                    }
                }
                (void)ld_acquire_u32(P.barrier);
            }
            __syncthreads();
            pt[5] += clock64() - t_bar;
            if (s_abort) { aborted = true; break; }
        }
    }
    cp_async_wait_all();
    if (P.prof && tid == 0) {
        pt[6] = clock64() - t_begin;
        for (int i = 0; i < 8; ++i) P.prof[(size_t)cta * 8 + i] = pt[i];
    }
}

// Decoder prologue: transposes the initial states into feature-major tiles, zeroes the <GO> frame, the context
// and the attention state (dirac at position 0 for monotonic attention, zeros for loc_sen).
struct DecInit {
    float *dst[4 + 4];
    const float *src[4 + 4];   // (N, width) row-major or null (= zeros)
    const float *vec[4 + 4];   // (width) one vector for every sentence (takes precedence), or null
    int width[4 + 4];
    int n;
    float *state0;
    int N, tiles, T_in, dirac;
};
__global__ void taco_dec_init_kernel(const DecInit d) {
    const size_t gt = blockIdx.x * (size_t)blockDim.x + threadIdx.x, gs = (size_t)gridDim.x * blockDim.x;
    for (int b = 0; b < d.n; ++b) {
        const size_t total = (size_t)d.tiles * d.width[b] * 32;
        for (size_t i = gt; i < total; i += gs) {
            const int r = (int)(i & 31);
            const size_t fk = i >> 5;
            const int tile = (int)(fk / d.width[b]), f = (int)(fk - (size_t)tile * d.width[b]);
            const int n = tile * 32 + r;
            d.dst[b][i] = d.vec[b] ? d.vec[b][f] : (d.src[b] && n < d.N) ? d.src[b][(size_t)n * d.width[b] + f] : 0.0f;
        }
    }
    for (size_t i = gt; i < (size_t)d.N * d.T_in; i += gs) d.state0[i] = (d.dirac && (i % d.T_in) == 0) ? 1.0f : 0.0f;
}

}  // namespace taco
