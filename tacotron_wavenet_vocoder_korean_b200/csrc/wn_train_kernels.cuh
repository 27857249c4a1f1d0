// wn_train_kernels.cuh -- hand-written sm_100a kernels of the WaveNet training step.
//
// The contractions of the step (dilated 2-tap convs, 1x1 convs and their data / weight gradients) are plain GEMMs and
// go to cuBLASLt (bf16 tensor cores, fp32 accumulation) from wn_train.cu; everything that is NOT a contraction is here:
//   upsample fwd/bwd (conv2d_transpose stack, model.py:102-111), causal conv fwd/bwd (scalar input: a 32-tap FIR into R
//   channels; one-hot input: two gathered kernel rows, model.py:41-46), gated activation fwd/bwd with the conditioning
//   biases (model.py:71-86), ReLU / bias / column-sum epilogues of the post-processing stack (model.py:150-165), the
//   discretized-mixture-of-logistics loss with its analytic gradient (mixture.py:27-81), the softmax cross-entropy head
//   of the one-hot model (model.py:292-296), global-condition gradients, L2, gradient norm, Adam + EMA (model.py:314-346).
//
// All activations are indexed by ABSOLUTE time: row = n*T0 + tau, tau = index of the causal-conv output.  Layer l
// (dilation d, input start s_l = sum of earlier dilations) produces valid outputs for tau >= off_l = s_l + d; rows
// before that hold finite junk in the forward pass and exact zeros in every gradient buffer, which is what lets each
// 2-tap conv / weight gradient run as ONE GEMM over all sentences (a tap is a row offset of the same matrix).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wnt {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float2 ld2(const float *p, size_t i) { return *reinterpret_cast<const float2 *>(p + i); }
__device__ __forceinline__ float2 ld2(const bf16 *p, size_t i) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p + i));
}
__device__ __forceinline__ void st2(float *p, size_t i, float2 v) { *reinterpret_cast<float2 *>(p + i) = v; }
__device__ __forceinline__ void st2(bf16 *p, size_t i, float2 v) {
    *reinterpret_cast<__nv_bfloat162 *>(p + i) = __float22bfloat162_rn(v);
}
__device__ __forceinline__ float ld1(const float *p, size_t i) { return p[i]; }
__device__ __forceinline__ float ld1(const bf16 *p, size_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void st1(float *p, size_t i, float v) { p[i] = v; }
__device__ __forceinline__ void st1(bf16 *p, size_t i, float v) { p[i] = __float2bfloat16_rn(v); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }

// Gate non-linearities: exact libdevice functions in fp32 mode; in bf16 mode the result is rounded to 8 mantissa bits
// anyway, so one MUFU.TANH each (tanh.approx.f32, |rel err| < 2^-10.9) -- sigmoid(x) = 0.5*tanh(x/2) + 0.5.
template <typename T> struct Act {
    static __device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
    static __device__ __forceinline__ float sigm_(float x) { return sigmoidf_(x); }
};
template <> struct Act<bf16> {
    static __device__ __forceinline__ float tanh_(float x) {
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float sigm_(float x) { return fmaf(0.5f, tanh_(0.5f * x), 0.5f); }
};

constexpr int EW_THREADS = 256;

// Column-pair tiling shared by the row-wise kernels: `cols` is a power of two in [2, 512]; thread -> (row lane, 2 columns).
struct ColTile {
    int tpr, rpp, lane_r, c;
    __device__ __forceinline__ ColTile(int cols) {
        tpr = cols >> 1;
        rpp = EW_THREADS / tpr;
        lane_r = threadIdx.x / tpr;
        c = (threadIdx.x - lane_r * tpr) * 2;
    }
};

// Sums (x, y) of the threads that share a column pair (same c, all row lanes) and atomically adds them to out[c], out[c+1].
__device__ __forceinline__ void block_colsum2(float2 v, const ColTile &t, float *out, float2 *red) {
    red[threadIdx.x] = v;
    __syncthreads();
    if (t.lane_r == 0) {
        float2 s = v;
        for (int r = 1; r < t.rpp; ++r) {
            const float2 o = red[r * t.tpr + threadIdx.x];
            s.x += o.x;
            s.y += o.y;
        }
        atomicAdd(out + t.c, s.x);
        atomicAdd(out + t.c + 1, s.y);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
// Upsampling stack.  One stage of conv2d_transpose(filters=1, kernel (F, 2), strides (F, 1), 'same', no bias):
//   out[n, i*F + a, w] = in[n, i, w]*K[a][0] + in[n, i, w-1]*K[a][1]            (oracle/np_oracle.py create_upsample)
// `rows_out` <= Ti*F truncates the last stage to the rows the network reads.
template <typename TO>
__global__ void ups_fwd_kernel(const float *__restrict__ in, const float *__restrict__ K, TO *__restrict__ out, int N, int Ti,
                               int F, int C, int rows_out) {
    const size_t total = (size_t)N * rows_out * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(e % C);
        const size_t r = e / C;
        const int j = (int)(r % rows_out), n = (int)(r / rows_out);
        const int i = j / F, a = j - i * F;
        const float *src = in + ((size_t)n * Ti + i) * C;
        float v = src[w] * K[2 * a];
        if (w > 0) v = fmaf(src[w - 1], K[2 * a + 1], v);
        st1(out, e, v);
    }
}

// Gradient of one stage: dK[a][0..1] (atomic, pre-zeroed) and, when din != null, the gradient of the stage input.
constexpr int UPS_MAX_F = 16;
__global__ void ups_bwd_kernel(const float *__restrict__ in, const float *__restrict__ K, const float *__restrict__ dout,
                               float *__restrict__ dK, float *__restrict__ din, int N, int Ti, int F, int C, int rows_out) {
    float s0[UPS_MAX_F], s1[UPS_MAX_F];
#pragma unroll
    for (int a = 0; a < UPS_MAX_F; ++a) s0[a] = s1[a] = 0.f;
    const size_t total = (size_t)N * Ti * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(e % C);
        const size_t r = e / C;
        const int i = (int)(r % Ti), n = (int)(r / Ti);
        const float x = in[e], xp = w > 0 ? in[e - 1] : 0.f;
        float g = 0.f;
#pragma unroll
        for (int a = 0; a < UPS_MAX_F; ++a) {
            if (a < F) {
                const int j = i * F + a;
                if (j < rows_out) {
                    const float *dr = dout + ((size_t)n * rows_out + j) * C;
                    const float d = dr[w];
                    s0[a] = fmaf(x, d, s0[a]);
                    s1[a] = fmaf(xp, d, s1[a]);
                    g = fmaf(d, K[2 * a], g);
                    if (w + 1 < C) g = fmaf(dr[w + 1], K[2 * a + 1], g);
                }
            }
        }
        if (din) din[e] = g;
    }
    __shared__ float red[2 * UPS_MAX_F][EW_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < UPS_MAX_F; ++a) {
        float u = s0[a], v = s1[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            u += __shfl_xor_sync(0xffffffffu, u, o);
            v += __shfl_xor_sync(0xffffffffu, v, o);
        }
        if (lane == 0) {
            red[2 * a][warp] = u;
            red[2 * a + 1][warp] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * F) {
        float s = 0.f;
        for (int wq = 0; wq < EW_THREADS / 32; ++wq) s += red[threadIdx.x][wq];
        atomicAdd(dK + threadIdx.x, s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Causal layer, scalar input (model.py:41-44): x0[n, tau, r] = sum_k wav[n, tau + k] * W[k][r], 'valid', no bias.
template <typename T>
__global__ void causal_fwd_kernel(const float *__restrict__ wav, const float *__restrict__ W, T *__restrict__ X0, int Tlen, int T0,
                                  int ifw, int R, int CH) {
    extern __shared__ float sm[];
    float *sW = sm;                 // ifw*R
    float *sx = sm + ifw * R;       // CH + ifw
    const int n = blockIdx.y, t0 = blockIdx.x * CH, t1 = min(T0, t0 + CH);
    for (int i = threadIdx.x; i < ifw * R; i += blockDim.x) sW[i] = W[i];
    for (int i = threadIdx.x; i < (t1 - t0) + ifw - 1; i += blockDim.x) sx[i] = wav[(size_t)n * Tlen + t0 + i];
    __syncthreads();
    const int tpr = R >> 1, rpp = blockDim.x / tpr, lr = threadIdx.x / tpr, c = (threadIdx.x - lr * tpr) * 2;
    for (int tau = t0 + lr; tau < t1; tau += rpp) {
        float2 acc = make_float2(0.f, 0.f);
        const float *xs = sx + (tau - t0);
        for (int k = 0; k < ifw; ++k) {
            const float x = xs[k];
            acc.x = fmaf(x, sW[k * R + c], acc.x);
            acc.y = fmaf(x, sW[k * R + c + 1], acc.y);
        }
        st2(X0, ((size_t)n * T0 + tau) * R + c, acc);
    }
}

// dW[k][r] = sum_{n,tau} wav[n, tau + k] * dX[n, tau, r]   (atomic into pre-zeroed dW).  Thread -> (r, tap group).
constexpr int CAUSAL_TAPS = 32;
template <typename TI>
__global__ void causal_bwd_kernel(const float *__restrict__ wav, const TI *__restrict__ dX, float *__restrict__ dW, int Tlen, int T0,
                                  int ifw, int R, int CH) {
    extern __shared__ float sx[];   // CH + ifw
    const int n = blockIdx.y, t0 = blockIdx.x * CH, t1 = min(T0, t0 + CH);
    for (int i = threadIdx.x; i < (t1 - t0) + ifw - 1; i += blockDim.x) sx[i] = wav[(size_t)n * Tlen + t0 + i];
    __syncthreads();
    const int KG = blockDim.x / R;                 // tap groups
    const int r = threadIdx.x % R, kg = threadIdx.x / R;
    const int per = (ifw + KG - 1) / KG;           // <= CAUSAL_TAPS (checked on the host)
    const int k0 = kg * per;
    float acc[CAUSAL_TAPS];
#pragma unroll
    for (int j = 0; j < CAUSAL_TAPS; ++j) acc[j] = 0.f;
    if (kg < KG) {
        for (int tau = t0; tau < t1; ++tau) {
            const float v = ld1(dX, ((size_t)n * T0 + tau) * R + r);
            const float *xs = sx + (tau - t0) + k0;
#pragma unroll
            for (int j = 0; j < CAUSAL_TAPS; ++j)
                if (j < per && k0 + j < ifw) acc[j] = fmaf(xs[j], v, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < CAUSAL_TAPS; ++j)
            if (j < per && k0 + j < ifw) atomicAdd(dW + (size_t)(k0 + j) * R + r, acc[j]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// One-hot input (scalar_input=False): add_loss mu-law encodes the waveform (wavenet/ops.py:22-33, evaluated in fp32 as the
// reference's graph does), one-hot encodes the ids (model.py:260) and the causal layer is a 'valid' conv of width
// filter_width = 2 over Q channels (model.py:41-46) -- i.e. two gathered kernel rows per step:
//   x0[n, tau, :] = W[0][id[n, tau]][:] + W[1][id[n, tau + 1]][:]
__global__ void mulaw_ids_kernel(const float *__restrict__ wav, int32_t *__restrict__ ids, size_t n, int Q) {
    const float mu = (float)(Q - 1), log_mu = log1pf(mu);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = wav[i];
        const float mag = __fdiv_rn(log1pf(__fmul_rn(mu, fminf(fabsf(x), 1.0f))), log_mu);
        const float sig = x > 0.f ? mag : (x < 0.f ? -mag : 0.f);
        ids[i] = (int32_t)__fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(sig, 1.0f), 0.5f), mu), 0.5f);
    }
}

template <typename T>
__global__ void onehot_causal_fwd_kernel(const int32_t *__restrict__ ids, const float *__restrict__ W, T *__restrict__ X0, int Tlen,
                                         int T0, long M, int Q, int R) {
    const ColTile t(R);
    for (long row = (long)blockIdx.x * t.rpp + t.lane_r; row < M; row += (long)gridDim.x * t.rpp) {
        const int n = (int)(row / T0), tau = (int)(row - (long)n * T0);
        const int a = ids[(size_t)n * Tlen + tau], b = ids[(size_t)n * Tlen + tau + 1];
        const float2 u = ld2(W, (size_t)a * R + t.c), v = ld2(W, ((size_t)Q + b) * R + t.c);
        st2(X0, (size_t)row * R + t.c, make_float2(u.x + v.x, u.y + v.y));
    }
}

// dW[0][id[n, tau]][:] += dX[n, tau, :], dW[1][id[n, tau + 1]][:] += dX[n, tau, :]   (atomic scatter into pre-zeroed dW)
template <typename TI>
__global__ void onehot_causal_bwd_kernel(const int32_t *__restrict__ ids, const TI *__restrict__ dX, float *__restrict__ dW, int Tlen,
                                         int T0, long M, int Q, int R) {
    const ColTile t(R);
    for (long row = (long)blockIdx.x * t.rpp + t.lane_r; row < M; row += (long)gridDim.x * t.rpp) {
        const int n = (int)(row / T0), tau = (int)(row - (long)n * T0);
        const int a = ids[(size_t)n * Tlen + tau], b = ids[(size_t)n * Tlen + tau + 1];
        const float2 g = ld2(dX, (size_t)row * R + t.c);
        float *pa = dW + (size_t)a * R + t.c, *pb = dW + ((size_t)Q + b) * R + t.c;
        atomicAdd(pa, g.x);
        atomicAdd(pa + 1, g.y);
        atomicAdd(pb, g.x);
        atomicAdd(pb + 1, g.y);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Gated activation (model.py:68-86).  FG = fp32 GEMM output [filter | gate] without biases; adds the layer bias and the
// per-sentence global-condition vector, stores tanh / sigmoid for the backward pass, z for the dense GEMM and -- for the
// last output_width steps -- z into this layer's column block of the concatenated skip operand.
template <typename T>
__global__ void gate_fwd_kernel(const float *__restrict__ FG, const float *__restrict__ bias, const float *__restrict__ gcb,
                                T *__restrict__ TS, T *__restrict__ Z, T *__restrict__ Zs, long row0, long M, int T0, int D, int SL,
                                int OW, int zs_col0, int LD) {
    const ColTile t(D);
    const int D2 = 2 * D;
    float2 bf = make_float2(0.f, 0.f), bg = bf;
    if (bias) {
        bf = make_float2(bias[t.c], bias[t.c + 1]);
        bg = make_float2(bias[D + t.c], bias[D + t.c + 1]);
    }
    for (long row = row0 + (long)blockIdx.x * t.rpp + t.lane_r; row < M; row += (long)gridDim.x * t.rpp) {
        const int n = (int)(row / T0), tau = (int)(row - (long)n * T0);
        float2 f = ld2(FG, (size_t)row * D2 + t.c), g = ld2(FG, (size_t)row * D2 + D + t.c);
        f.x += bf.x; f.y += bf.y; g.x += bg.x; g.y += bg.y;
        if (gcb) {
            const float *gb = gcb + (size_t)n * D2;
            f.x += gb[t.c]; f.y += gb[t.c + 1]; g.x += gb[D + t.c]; g.y += gb[D + t.c + 1];
        }
        const float2 th = make_float2(Act<T>::tanh_(f.x), Act<T>::tanh_(f.y));
        const float2 sg = make_float2(Act<T>::sigm_(g.x), Act<T>::sigm_(g.y));
        st2(TS, (size_t)row * D2 + t.c, th);
        st2(TS, (size_t)row * D2 + D + t.c, sg);
        const float2 z = make_float2(th.x * sg.x, th.y * sg.y);
        st2(Z, (size_t)row * D + t.c, z);
        if (tau >= SL) st2(Zs, ((size_t)n * OW + (tau - SL)) * LD + zs_col0 + t.c, z);
    }
}

// Backward of the gate: dz = dZ32 (dense path, may be null) + dZs (skip path, last OW steps) ->
//   dFG = [dz*sg*(1-th^2) | dz*th*sg*(1-sg)], exact zeros for tau < off_l; Z = th*sg re-materialised for the dense weight
//   gradient (may be null); per-sentence column sums of dFG into SB (n, 2D) for the bias and global-condition gradients.
template <typename T>
__global__ void gate_bwd_kernel(const float *__restrict__ dZ32, const T *__restrict__ dZs, const T *__restrict__ TS, T *__restrict__ dFG,
                                T *__restrict__ Z, float *__restrict__ SB, long row0, int T0, int D, int SL, int OW, int off_l,
                                int zs_col0, int LD, int CH) {
    __shared__ float2 red[EW_THREADS];
    const ColTile t(D);
    const int D2 = 2 * D;
    const int n = blockIdx.y, t0 = blockIdx.x * CH, t1 = min(T0, t0 + CH);
    float2 sf = make_float2(0.f, 0.f), sgs = sf;
    const float2 zero = make_float2(0.f, 0.f);
    for (int tau = t0 + t.lane_r; tau < t1; tau += t.rpp) {
        const long row = (long)n * T0 + tau;
        if (row < row0) continue;
        if (tau < off_l) {
            st2(dFG, (size_t)row * D2 + t.c, zero);
            st2(dFG, (size_t)row * D2 + D + t.c, zero);
            if (Z) st2(Z, (size_t)row * D + t.c, zero);
            continue;
        }
        float2 dz = dZ32 ? ld2(dZ32, (size_t)row * D + t.c) : zero;
        if (tau >= SL) {
            const float2 s = ld2(dZs, ((size_t)n * OW + (tau - SL)) * LD + zs_col0 + t.c);
            dz.x += s.x; dz.y += s.y;
        }
        const float2 th = ld2(TS, (size_t)row * D2 + t.c), sg = ld2(TS, (size_t)row * D2 + D + t.c);
        const float2 df = make_float2(dz.x * sg.x * (1.f - th.x * th.x), dz.y * sg.y * (1.f - th.y * th.y));
        const float2 dg = make_float2(dz.x * th.x * sg.x * (1.f - sg.x), dz.y * th.y * sg.y * (1.f - sg.y));
        st2(dFG, (size_t)row * D2 + t.c, df);
        st2(dFG, (size_t)row * D2 + D + t.c, dg);
        if (Z) st2(Z, (size_t)row * D + t.c, make_float2(th.x * sg.x, th.y * sg.y));
        sf.x += df.x; sf.y += df.y; sgs.x += dg.x; sgs.y += dg.y;
    }
    block_colsum2(sf, t, SB + (size_t)n * D2, red);
    block_colsum2(sgs, t, SB + (size_t)n * D2 + D, red);
}

// out = T(in) over rows [row0, M) and column sums of the same rows into colsum (atomic, pre-zeroed; may be null).
template <typename T>
__global__ void cast_colsum_kernel(const float *__restrict__ in, T *__restrict__ out, float *__restrict__ colsum, long row0, long M,
                                   int cols) {
    __shared__ float2 red[EW_THREADS];
    const ColTile t(cols);
    float2 s = make_float2(0.f, 0.f);
    for (long row = row0 + (long)blockIdx.x * t.rpp + t.lane_r; row < M; row += (long)gridDim.x * t.rpp) {
        const float2 v = ld2(in, (size_t)row * cols + t.c);
        st2(out, (size_t)row * cols + t.c, v);
        s.x += v.x; s.y += v.y;
    }
    if (colsum) block_colsum2(s, t, colsum, red);
}

// out = relu(in + bias) (model.py:157-158: transformed1 = relu(sum of skips), the summed skip biases come in as `bias`).
template <typename T>
__global__ void bias_relu_kernel(const float *__restrict__ in, const float *__restrict__ bias, T *__restrict__ out, long M, int cols) {
    const ColTile t(cols);
    float2 b = make_float2(0.f, 0.f);
    if (bias) b = make_float2(bias[t.c], bias[t.c + 1]);
    for (long row = (long)blockIdx.x * t.rpp + t.lane_r; row < M; row += (long)gridDim.x * t.rpp) {
        float2 v = ld2(in, (size_t)row * cols + t.c);
        v.x = fmaxf(v.x + b.x, 0.f);
        v.y = fmaxf(v.y + b.y, 0.f);
        st2(out, (size_t)row * cols + t.c, v);
    }
}

// out = d * (act > 0), column sums of out into dbias (atomic, pre-zeroed; may be null).
template <typename T>
__global__ void relu_bwd_colsum_kernel(const float *__restrict__ d, const T *__restrict__ act, T *__restrict__ out,
                                       float *__restrict__ dbias, long M, int cols) {
    __shared__ float2 red[EW_THREADS];
    const ColTile t(cols);
    float2 s = make_float2(0.f, 0.f);
    for (long row = (long)blockIdx.x * t.rpp + t.lane_r; row < M; row += (long)gridDim.x * t.rpp) {
        float2 v = ld2(d, (size_t)row * cols + t.c);
        const float2 a = ld2(act, (size_t)row * cols + t.c);
        v.x = a.x > 0.f ? v.x : 0.f;
        v.y = a.y > 0.f ? v.y : 0.f;
        st2(out, (size_t)row * cols + t.c, v);
        s.x += v.x; s.y += v.y;
    }
    if (dbias) block_colsum2(s, t, dbias, red);
}

// bsum[c] = sum_l BS[l][c]
__global__ void skip_bias_sum_kernel(const float *__restrict__ BS, float *__restrict__ bsum, int L, int S) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S) return;
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += BS[(size_t)l * S + c];
    bsum[c] = s;
}
// every layer's skip/bias gradient is the same vector
__global__ void skip_bias_bcast_kernel(const float *__restrict__ g, float *__restrict__ dBS, int L, int S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L * S) dBS[i] = g[i % S];
}

// ---------------------------------------------------------------------------------------------------------------------
// discretized_mix_logistic_loss(num_class = 65536, reduce=False) (mixture.py:27-81) + reduce_mean (model.py:290) and its
// gradient w.r.t. the network output.  One thread per (n, t) row of Y (Mo, OP) fp32; dY (Mo, OP) = d mean / dY; column
// sums of dY into dbias (the conv1d_2/bias gradient); the loss sum in double into acc[0].
constexpr int MOL_MAX_K = 16;
template <typename T>
__global__ void mol_loss_kernel(const float *__restrict__ Y, const float *__restrict__ wav, T *__restrict__ dY, float *__restrict__ dbias,
                                double *__restrict__ acc, long Mo, int OW, int Tlen, int rf, int K, int OP, float log_scale_min,
                                float half_bin, float log_half_classes, float inv_count) {
    __shared__ float sdb[EW_THREADS / 32][3 * MOL_MAX_K];
    __shared__ float sloss[EW_THREADS / 32];
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float loss = 0.f;
    float g[3 * MOL_MAX_K];
#pragma unroll
    for (int i = 0; i < 3 * MOL_MAX_K; ++i) g[i] = 0.f;
    if (row < Mo) {
        const int n = (int)(row / OW), j = (int)(row - (long)n * OW);
        const float y = wav[(size_t)n * Tlen + rf + j];
        const float *yr = Y + (size_t)row * OP;
        float lg[MOL_MAX_K], lp[MOL_MAX_K], dmu[MOL_MAX_K], dls[MOL_MAX_K];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < MOL_MAX_K; ++k)
            if (k < K) {
                lg[k] = yr[k];
                mx = fmaxf(mx, lg[k]);
            }
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < MOL_MAX_K; ++k)
            if (k < K) se += expf(lg[k] - mx);
        const float lse_logits = mx + logf(se);
        float amax = -INFINITY;
#pragma unroll
        for (int k = 0; k < MOL_MAX_K; ++k)
            if (k < K) {
                const float mu = yr[K + k], ls_raw = yr[2 * K + k];
                const float ls = fmaxf(ls_raw, log_scale_min);
                const float c = y - mu, inv = expf(-ls);
                const float pin = inv * (c + half_bin), min_ = inv * (c - half_bin), mid = inv * c;
                float logp, dm, ds;
                if (y < -0.999f) {                                   // log cdf_plus
                    logp = pin - softplusf_(pin);
                    const float d = 1.f - sigmoidf_(pin);
                    dm = -inv * d;
                    ds = -pin * d;
                } else if (y > 0.999f) {                             // log (1 - cdf_min)
                    logp = -softplusf_(min_);
                    const float d = -sigmoidf_(min_);
                    dm = -inv * d;
                    ds = -min_ * d;
                } else {
                    const float cp = sigmoidf_(pin), cm = sigmoidf_(min_);
                    const float delta = cp - cm;
                    if (delta > 1e-5f) {
                        logp = logf(fmaxf(delta, 1e-12f));
                        const float dp = cp * (1.f - cp) / delta, dn = -cm * (1.f - cm) / delta;
                        dm = -inv * (dp + dn);
                        ds = -(pin * dp + min_ * dn);
                    } else {                                         // log pdf at the bin centre
                        logp = mid - ls - 2.f * softplusf_(mid) - log_half_classes;
                        const float d = 1.f - 2.f * sigmoidf_(mid);
                        dm = -inv * d;
                        ds = -mid * d - 1.f;
                    }
                }
                if (!(ls_raw >= log_scale_min)) ds = 0.f;            // tf.maximum passes the gradient where x >= y
                dmu[k] = dm;
                dls[k] = ds;
                lp[k] = logp + (lg[k] - lse_logits);
                amax = fmaxf(amax, lp[k]);
            }
        float sa = 0.f;
#pragma unroll
        for (int k = 0; k < MOL_MAX_K; ++k)
            if (k < K) sa += expf(lp[k] - amax);
        const float lse = amax + logf(sa);
        loss = -lse;
#pragma unroll
        for (int k = 0; k < MOL_MAX_K; ++k)
            if (k < K) {
                const float w = expf(lp[k] - lse);                   // responsibility of component k
                const float pi = expf(lg[k] - lse_logits);
                g[k] = (pi - w) * inv_count;
                g[MOL_MAX_K + k] = -w * dmu[k] * inv_count;
                g[2 * MOL_MAX_K + k] = -w * dls[k] * inv_count;
            }
        T *dr = dY + (size_t)row * OP;
#pragma unroll
        for (int k = 0; k < MOL_MAX_K; ++k)
            if (k < K) {
                st1(dr, k, g[k]);
                st1(dr, K + k, g[MOL_MAX_K + k]);
                st1(dr, 2 * K + k, g[2 * MOL_MAX_K + k]);
            }
        for (int c = 3 * K; c < OP; ++c) st1(dr, c, 0.f);
    }
    // block reductions: loss and the 3K bias-gradient columns
#pragma unroll
    for (int i = 0; i < 3 * MOL_MAX_K; ++i) {
        float v = g[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sdb[warp][i] = v;
    }
    float lv = loss;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lv += __shfl_xor_sync(0xffffffffu, lv, o);
    if (lane == 0) sloss[warp] = lv;
    __syncthreads();
    if (threadIdx.x < 3 * MOL_MAX_K) {
        const int comp = threadIdx.x / MOL_MAX_K, k = threadIdx.x % MOL_MAX_K;
        if (k < K && dbias) {
            float s = 0.f;
            for (int w = 0; w < EW_THREADS / 32; ++w) s += sdb[w][threadIdx.x];
            atomicAdd(dbias + comp * K + k, s);
        }
    }
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < EW_THREADS / 32; ++w) s += (double)sloss[w];
        atomicAdd(acc, s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// softmax_cross_entropy_with_logits_v2 + reduce_mean (model.py:292-296) against the one-hot target id[n, rf + j], and its
// gradient (softmax - one_hot) / count.  One warp per row of Y (Mo, OP) fp32, Q <= 32 * CE_PER_LANE classes; column sums of
// dY (the conv1d_2/bias gradient) are kept per lane across the warp's rows, merged in shared memory, one atomic per column
// and block; the loss sum in double into acc[0].
constexpr int CE_PER_LANE = 16;
template <typename T>
__global__ void softmax_ce_kernel(const float *__restrict__ Y, const int32_t *__restrict__ ids, T *__restrict__ dY,
                                  float *__restrict__ dbias, double *__restrict__ acc, long Mo, int OW, int Tlen, int rf, int Q, int OP,
                                  float inv_count) {
    __shared__ float scol[32 * CE_PER_LANE];
    __shared__ float sloss[EW_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 32 * CE_PER_LANE; i += blockDim.x) scol[i] = 0.f;
    __syncthreads();
    float cs[CE_PER_LANE];
#pragma unroll
    for (int i = 0; i < CE_PER_LANE; ++i) cs[i] = 0.f;
    float loss = 0.f;
    for (long row = (long)blockIdx.x * nw + warp; row < Mo; row += (long)gridDim.x * nw) {
        const int n = (int)(row / OW), j = (int)(row - (long)n * OW);
        const int tgt = ids[(size_t)n * Tlen + rf + j];
        const float *yr = Y + (size_t)row * OP;
        float v[CE_PER_LANE];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < CE_PER_LANE; ++i) {
            const int q = lane + 32 * i;
            v[i] = q < Q ? yr[q] : -INFINITY;
            mx = fmaxf(mx, v[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
#pragma unroll
        for (int i = 0; i < CE_PER_LANE; ++i)
            if (lane + 32 * i < Q) se += expf(v[i] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        const float lse = mx + logf(se);
        if (lane == 0) loss += lse - yr[tgt];
        T *dr = dY + (size_t)row * OP;
#pragma unroll
        for (int i = 0; i < CE_PER_LANE; ++i) {
            const int q = lane + 32 * i;
            if (q < Q) {
                const float g = (expf(v[i] - lse) - (q == tgt ? 1.f : 0.f)) * inv_count;
                st1(dr, q, g);
                cs[i] += g;
            }
        }
        for (int c = Q + lane; c < OP; c += 32) st1(dr, c, 0.f);
    }
#pragma unroll
    for (int i = 0; i < CE_PER_LANE; ++i)
        if (lane + 32 * i < Q) atomicAdd(&scol[lane + 32 * i], cs[i]);
    if (lane == 0) sloss[warp] = loss;
    __syncthreads();
    if (dbias)
        for (int i = threadIdx.x; i < Q; i += blockDim.x) atomicAdd(dbias + i, scol[i]);
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += (double)sloss[w];
        atomicAdd(acc, s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Global conditioning (model.py:71-73, 181-212).  GCB[l][n][:] = E[id_n] @ WGC_l  (G x 2D per layer, layer stride ls).
__global__ void gc_bias_kernel(const float *__restrict__ E, const int32_t *__restrict__ ids, const float *__restrict__ WGC, size_t ls,
                               float *__restrict__ GCB, int N, int G, int D2) {
    const int l = blockIdx.x, n = blockIdx.y;
    const float *e = E + (size_t)ids[n] * G;
    const float *w = WGC + (size_t)l * ls;
    for (int c = threadIdx.x; c < D2; c += blockDim.x) {
        float s = 0.f;
        for (int g = 0; g < G; ++g) s = fmaf(e[g], w[(size_t)g * D2 + c], s);
        GCB[((size_t)l * N + n) * D2 + c] = s;
    }
}
// dBFG[l][c] = sum_n SB[l][n][c]  (bias layer stride lsb)
__global__ void sb_reduce_kernel(const float *__restrict__ SB, float *__restrict__ dB, size_t lsb, int N, int D2) {
    const int l = blockIdx.x;
    for (int c = threadIdx.x; c < D2; c += blockDim.x) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += SB[((size_t)l * N + n) * D2 + c];
        dB[(size_t)l * lsb + c] = s;
    }
}
// dWGC[l][g][c] = sum_n E[id_n][g] * SB[l][n][c]
__global__ void gc_wgrad_kernel(const float *__restrict__ E, const int32_t *__restrict__ ids, const float *__restrict__ SB,
                                float *__restrict__ dWGC, size_t ls, int N, int G, int D2) {
    const int l = blockIdx.x, g = blockIdx.y;
    for (int c = threadIdx.x; c < D2; c += blockDim.x) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(E[(size_t)ids[n] * G + g], SB[((size_t)l * N + n) * D2 + c], s);
        dWGC[(size_t)l * ls + (size_t)g * D2 + c] = s;
    }
}
// dE[id_n][g] += sum_l sum_c SB[l][n][c] * WGC_l[g][c]   (atomic: sentences share speakers; dE pre-zeroed)
__global__ void gc_egrad_kernel(const int32_t *__restrict__ ids, const float *__restrict__ SB, const float *__restrict__ WGC, size_t ls,
                                float *__restrict__ dE, int L, int N, int G, int D2) {
    __shared__ float red[EW_THREADS / 32];
    const int n = blockIdx.x, g = blockIdx.y;
    float s = 0.f;
    for (int i = threadIdx.x; i < L * D2; i += blockDim.x) {
        const int l = i / D2, c = i - l * D2;
        s = fmaf(SB[((size_t)l * N + n) * D2 + c], WGC[(size_t)l * ls + (size_t)g * D2 + c], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
        atomicAdd(dE + (size_t)ids[n] * G + g, tot);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Optimizer side (model.py:303-308, 314-346).
// L2: g += s*p over the kernels (first n floats of the flat buffer), acc[1] += s/2 * sum p^2.
__global__ void l2_kernel(const float *__restrict__ p, float *__restrict__ g, double *__restrict__ acc, size_t n, float s) {
    double loc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = p[i];
        g[i] = fmaf(s, v, g[i]);
        loc += (double)v * v;
    }
    __shared__ double red[EW_THREADS];
    red[threadIdx.x] = loc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(acc + 1, 0.5 * (double)s * red[0]);
}
__global__ void sumsq_kernel(const float *__restrict__ g, double *__restrict__ out, size_t n) {
    double loc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) loc += (double)g[i] * g[i];
    __shared__ double red[EW_THREADS];
    red[threadIdx.x] = loc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, red[0]);
}
__global__ void loss_finish_kernel(const double *__restrict__ acc, float *__restrict__ loss, double inv_count) {
    loss[0] = (float)(acc[0] * inv_count + acc[1]);
}

// Adam (tf.train.AdamOptimizer: p -= lr_t * m / (sqrt(v) + eps), lr_t precomputed) + EMA shadow + compute-dtype copy.
// sumsq (may be null): squared gradient norm BEFORE grad_scale for tf.clip_by_global_norm(grads, clip_norm).
template <typename T>
__global__ void adam_ema_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                                float *__restrict__ ema, T *__restrict__ pc, size_t n, float lr_t, float b1, float b2, float eps,
                                float ema_decay, float grad_scale, const double *__restrict__ sumsq, float clip_norm) {
    float scale = grad_scale;
    if (sumsq) {
        const float gn = (float)sqrt(sumsq[0]) * grad_scale;
        scale *= clip_norm / fmaxf(gn, clip_norm);
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * scale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        const float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
        m[i] = mi;
        v[i] = vi;
        p[i] = pi;
        const float e = ema[i];
        ema[i] = e - (1.f - ema_decay) * (e - pi);
        if (pc) st1(pc, i, pi);
    }
}
template <typename T>
__global__ void cast_kernel(const float *__restrict__ in, T *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) st1(out, i, in[i]);
}

}  // namespace wnt
