// wn_mel.cuh -- STFT -> mel -> dB -> normalise (the reference's utils/audio.py:69-75 melspectrogram chain:
// preemphasis :22-25, librosa.stft(n_fft, hop, win) :139-143, mel basis dot :181-199, _amp_to_db :201-203,
// _normalize :208-212) as one sm_100a kernel: one CTA iteration per pair of frames.
//
// librosa evaluates the FFT in float64 (the pre-emphasised signal is float64) and stores complex64; a float32
// FFT would miss the 1e-4 tolerance in bands far below the frame's peak (absolute error ~1e-7 * peak), so the
// FFT runs in fp64 (Stockham radix-8 passes in registers + shared memory, see below; ~45 k fp64 operations per frame pair),
// components are rounded to fp32 like complex64, and everything after is fp32 like numpy.
// Included by wn_api.cu (single translation unit).
#pragma once

struct WnMelParams {
    const float *wav;          // [rows][n]
    long long n;               // samples per row
    int rows, frames;          // frames = 1 + n / hop
    int n_fft, log2n, hop, win, win_off;   // win_off = (n_fft - win) / 2
    int n_mels, n_bins;        // n_bins = n_fft/2 + 1
    const double *window;      // [win]  periodic Hann
    const double2 *twiddle;    // [n_fft/2]  exp(-2 pi i k / n_fft)
    const int *mel_start;      // [n_mels] first non-zero bin
    const int *mel_len;        // [n_mels]
    const int *mel_off;        // [n_mels] offset into mel_w
    const float *mel_w;        // packed non-zero filter weights
    double preemph;            // 0 disables
    float min_level, ref_level_db, min_level_db, max_abs;
    float *out;                // [rows][frames][n_mels]
};

// ---- Stockham radix-8 kernel (n_fft <= 4096): the path wn_melspectrogram takes ------------------------------------------------
// The radix-2 kernel below moves every point through shared memory 11 times (2048 points) with a block barrier per stage and is
// bound by exactly that (shared-memory bandwidth + barriers, not by the fp64 pipe).  Here a thread holds 8 points in registers,
// applies the inter-pass twiddles, runs an 8-point DFT (three radix-2 levels, no shared memory) and scatters the results into the
// other half of a ping-pong buffer: 2048 points = 8 * 8 * 8 * 4 -> FOUR passes, one barrier each, autosorting (no bit reversal).
// Pass with p = product of the previous radices, r = radix, t = N / r butterflies:
//   i in [0, t): k = i & (p - 1), j = (i - k) * r + k;  u[m] = x[i + m t] * exp(-2 pi i k m / (p r));  y[j + m p] = DFT_r(u)[m].
// Element e lives at e + (e >> 3) (one pad per 8 double2): the first pass stores with a stride of 8 elements.
struct cd { double x, y; };
__device__ __forceinline__ cd cadd(cd a, cd b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cd csub(cd a, cd b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cd cmul(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cmul_mi(cd a) { return {a.y, -a.x}; }                 // a * (-i)
__device__ __forceinline__ int mel_pad(int e) { return e + (e >> 3); }

__device__ __forceinline__ void mel_dft8(cd *a)
{
    const double S = 0.70710678118654752440;
    cd b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]), b1 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
    cd b2 = cadd(a[2], a[6]), b6 = csub(a[2], a[6]), b3 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
    b5 = {(b5.x + b5.y) * S, (b5.y - b5.x) * S};                                   // * exp(-i pi / 4)
    b6 = cmul_mi(b6);
    b7 = {(b7.y - b7.x) * S, -(b7.y + b7.x) * S};                                  // * exp(-3 i pi / 4)
    const cd c0 = cadd(b0, b2), c2 = csub(b0, b2), c1 = cadd(b1, b3), c3 = cmul_mi(csub(b1, b3));
    const cd c4 = cadd(b4, b6), c6 = csub(b4, b6), c5 = cadd(b5, b7), c7 = cmul_mi(csub(b5, b7));
    a[0] = cadd(c0, c1); a[4] = csub(c0, c1); a[2] = cadd(c2, c3); a[6] = csub(c2, c3);
    a[1] = cadd(c4, c5); a[5] = csub(c4, c5); a[3] = cadd(c6, c7); a[7] = csub(c6, c7);
}
__device__ __forceinline__ void mel_dft4(cd *a)
{
    const cd b0 = cadd(a[0], a[2]), b2 = csub(a[0], a[2]), b1 = cadd(a[1], a[3]), b3 = cmul_mi(csub(a[1], a[3]));
    a[0] = cadd(b0, b1); a[1] = cadd(b2, b3); a[2] = csub(b0, b1); a[3] = csub(b2, b3);
}

template <int RADIX>
__device__ __forceinline__ void mel_pass(const double2 *__restrict__ x, double2 *__restrict__ y, const double2 *__restrict__ tw, int n, int p, int tid, int nth)
{
    const int t = n / RADIX;
    const int step = n / (p * RADIX);            // twiddle exp(-2 pi i k m / (p r)) = table[k m step], table has n / 2 entries
    for (int i = tid; i < t; i += nth) {
        const int k = i & (p - 1), j = (i - k) * RADIX + k;
        cd u[RADIX];
#pragma unroll
        for (int m = 0; m < RADIX; ++m) {
            const double2 v = x[mel_pad(i + m * t)];
            u[m] = {v.x, v.y};
        }
        if (p > 1) {
#pragma unroll
            for (int m = 1; m < RADIX; ++m) {
                int idx = k * m * step;
                const bool neg = idx >= (n >> 1);
                if (neg) idx -= n >> 1;
                const double2 w = __ldg(tw + idx);
                const cd wv = neg ? cd{-w.x, -w.y} : cd{w.x, w.y};
                u[m] = cmul(u[m], wv);
            }
        }
        if (RADIX == 8) mel_dft8(u);
        else if (RADIX == 4) mel_dft4(u);
        else { const cd s0 = cadd(u[0], u[1]), s1 = csub(u[0], u[1]); u[0] = s0; u[1] = s1; }
#pragma unroll
        for (int m = 0; m < RADIX; ++m) y[mel_pad(j + m * p)] = make_double2(u[m].x, u[m].y);
    }
}

// CTAS = resident CTAs per SM the register allocation is bounded for: 2 -> 122 registers, 3 -> 80 registers (24 bytes of spill)
template <int CTAS>
__global__ void __launch_bounds__(256, CTAS) wn_mel_kernel_s8(const WnMelParams p)
{
    extern __shared__ __align__(16) unsigned char mel_smem[];
    const int padded = p.n_fft + (p.n_fft >> 3);
    double2 *bufA = reinterpret_cast<double2 *>(mel_smem);                 // [padded]
    double2 *bufB = bufA + padded;                                         // [padded]
    const int tid = threadIdx.x;
    const int nth = blockDim.x;
    const int pairs = (p.frames + 1) / 2;
    const int n8 = p.log2n / 3, rem = p.log2n - 3 * n8;                    // n_fft = 8^n8 * 2^rem
    for (long long item = blockIdx.x; item < (long long)p.rows * pairs; item += gridDim.x) {
        const int row = (int)(item / pairs), f0 = 2 * (int)(item % pairs);
        const bool two = f0 + 1 < p.frames;
        const float *x = p.wav + (size_t)row * p.n;
        // 1. windowed, pre-emphasised, reflect-padded frames f0 (real part) and f0 + 1 (imaginary part), natural order
        for (int i = tid; i < p.n_fft; i += nth) {
            double v[2] = {0.0, 0.0};
            const int wi = i - p.win_off;
            if (wi >= 0 && wi < p.win) {
                const double wv = p.window[wi];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (u == 1 && !two) break;
                    long long q = (long long)(f0 + u) * p.hop + i - p.n_fft / 2;
                    if (q < 0) q = -q;
                    if (q >= p.n) q = 2 * (p.n - 1) - q;
                    double y = (double)x[q];
                    if (p.preemph != 0.0 && q > 0) y = __dadd_rn(y, __dmul_rn(-p.preemph, (double)x[q - 1]));
                    v[u] = __dmul_rn(wv, y);
                }
            }
            bufA[mel_pad(i)] = make_double2(v[0], v[1]);
        }
        __syncthreads();
        // 2. Stockham passes, fp64
        double2 *src = bufA, *dst = bufB;
        int pp = 1;
        for (int s = 0; s < n8; ++s) {
            mel_pass<8>(src, dst, p.twiddle, p.n_fft, pp, tid, nth);
            __syncthreads();
            double2 *sw = src; src = dst; dst = sw;
            pp *= 8;
        }
        if (rem == 2) mel_pass<4>(src, dst, p.twiddle, p.n_fft, pp, tid, nth);
        else if (rem == 1) mel_pass<2>(src, dst, p.twiddle, p.n_fft, pp, tid, nth);
        if (rem) {
            __syncthreads();
            double2 *sw = src; src = dst; dst = sw;
        }
        // 3. split the two spectra; |D| with the components rounded to fp32 first (complex64), np.abs -> hypot.  The spectrum is in
        //    `src`; the magnitudes go to the other buffer.
        float *mag = reinterpret_cast<float *>(dst);                        // [2][n_bins]
        for (int k = tid; k < p.n_bins; k += nth) {
            const double2 zk = src[mel_pad(k)], zn = src[mel_pad((p.n_fft - k) & (p.n_fft - 1))];
            const double re1 = (double)(float)(0.5 * (zk.x + zn.x)), im1 = (double)(float)(0.5 * (zk.y - zn.y));
            const double re2 = (double)(float)(0.5 * (zk.y + zn.y)), im2 = (double)(float)(0.5 * (zn.x - zk.x));
            mag[k] = (float)sqrt(re1 * re1 + im1 * im1);
            mag[p.n_bins + k] = (float)sqrt(re2 * re2 + im2 * im2);
        }
        __syncthreads();
        // 4. mel filterbank (sparse triangles), dB, reference level, symmetric normalisation + clip; all fp32
        for (int o = tid; o < 2 * p.n_mels; o += nth) {
            const int u = o / p.n_mels, c = o - u * p.n_mels;
            if (u == 1 && !two) continue;
            const float *w = p.mel_w + p.mel_off[c];
            const float *m = mag + u * p.n_bins + p.mel_start[c];
            float acc = 0.0f;
            for (int j = 0; j < p.mel_len[c]; ++j) acc = __fmaf_rn(w[j], m[j], acc);
            float S = __fsub_rn(__fmul_rn(20.0f, log10f(fmaxf(p.min_level, acc))), p.ref_level_db);
            float v = __fsub_rn(__fmul_rn(2.0f * p.max_abs, __fdiv_rn(__fsub_rn(S, p.min_level_db), -p.min_level_db)), p.max_abs);
            v = fminf(fmaxf(v, -p.max_abs), p.max_abs);
            p.out[((size_t)row * p.frames + f0 + u) * p.n_mels + c] = v;
        }
        __syncthreads();
    }
}

// ---- radix-2 kernel: n_fft = 8192 only (its ping-pong buffers would not fit), and the A/B partner of the kernel above ------------
// Two frames per FFT: frames f and f+1 of a row are the real and imaginary parts of one complex 2048-point transform,
// z = x1 + i*x2  ->  X1[k] = (Z[k] + conj(Z[N-k])) / 2,  X2[k] = (Z[k] - conj(Z[N-k])) / (2i)  (exact up to fp64 rounding),
// which halves the butterfly work and the number of block barriers per frame.
extern "C" __global__ void __launch_bounds__(256) wn_mel_kernel(const WnMelParams p)
{
    extern __shared__ __align__(16) unsigned char mel_smem[];
    double2 *buf = reinterpret_cast<double2 *>(mel_smem);                  // [n_fft]
    float *mag = reinterpret_cast<float *>(buf + p.n_fft);                 // [2][n_bins]
    const int tid = threadIdx.x;
    const int nth = blockDim.x;
    const int pairs = (p.frames + 1) / 2;
    for (long long item = blockIdx.x; item < (long long)p.rows * pairs; item += gridDim.x) {
        const int row = (int)(item / pairs), f0 = 2 * (int)(item % pairs);
        const bool two = f0 + 1 < p.frames;
        const float *x = p.wav + (size_t)row * p.n;
        // 1. windowed, pre-emphasised, reflect-padded frames, stored in bit-reversed order
        for (int i = tid; i < p.n_fft; i += nth) {
            double v[2] = {0.0, 0.0};
            const int wi = i - p.win_off;
            if (wi >= 0 && wi < p.win) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (u == 1 && !two) break;
                    long long q = (long long)(f0 + u) * p.hop + i - p.n_fft / 2;
                    if (q < 0) q = -q;
                    if (q >= p.n) q = 2 * (p.n - 1) - q;
                    double y = (double)x[q];
                    if (p.preemph != 0.0 && q > 0) y = __dadd_rn(y, __dmul_rn(-p.preemph, (double)x[q - 1]));
                    v[u] = __dmul_rn(p.window[wi], y);
                }
            }
            const unsigned r = __brev((unsigned)i) >> (32 - p.log2n);
            buf[r] = make_double2(v[0], v[1]);
        }
        __syncthreads();
        // 2. radix-2 decimation-in-time FFT, fp64
        for (int s = 1; s <= p.log2n; ++s) {
            const int half = 1 << (s - 1);
            const int tw_stride = (p.n_fft >> 1) >> (s - 1);
            for (int j = tid; j < (p.n_fft >> 1); j += nth) {
                const int pos = j & (half - 1);
                const int i0 = ((j >> (s - 1)) << s) + pos, i1 = i0 + half;
                const double2 w = p.twiddle[pos * tw_stride];
                const double2 a = buf[i0], b = buf[i1];
                const double tr = w.x * b.x - w.y * b.y, ti = w.x * b.y + w.y * b.x;
                buf[i0] = make_double2(a.x + tr, a.y + ti);
                buf[i1] = make_double2(a.x - tr, a.y - ti);
            }
            __syncthreads();
        }
        // 3. split the two spectra; |D| with the components rounded to fp32 first (complex64), np.abs -> hypot
        for (int k = tid; k < p.n_bins; k += nth) {
            const double2 zk = buf[k], zn = buf[(p.n_fft - k) & (p.n_fft - 1)];
            const double re1 = (double)(float)(0.5 * (zk.x + zn.x)), im1 = (double)(float)(0.5 * (zk.y - zn.y));
            const double re2 = (double)(float)(0.5 * (zk.y + zn.y)), im2 = (double)(float)(0.5 * (zn.x - zk.x));
            mag[k] = (float)sqrt(re1 * re1 + im1 * im1);
            mag[p.n_bins + k] = (float)sqrt(re2 * re2 + im2 * im2);
        }
        __syncthreads();
        // 4. mel filterbank (sparse triangles), dB, reference level, symmetric normalisation + clip; all fp32
        for (int o = tid; o < 2 * p.n_mels; o += nth) {
            const int u = o / p.n_mels, c = o - u * p.n_mels;
            if (u == 1 && !two) continue;
            const float *w = p.mel_w + p.mel_off[c];
            const float *m = mag + u * p.n_bins + p.mel_start[c];
            float acc = 0.0f;
            for (int j = 0; j < p.mel_len[c]; ++j) acc = __fmaf_rn(w[j], m[j], acc);
            float S = __fsub_rn(__fmul_rn(20.0f, log10f(fmaxf(p.min_level, acc))), p.ref_level_db);
            float v = __fsub_rn(__fmul_rn(2.0f * p.max_abs, __fdiv_rn(__fsub_rn(S, p.min_level_db), -p.min_level_db)), p.max_abs);
            v = fminf(fmaxf(v, -p.max_abs), p.max_abs);
            p.out[((size_t)row * p.frames + f0 + u) * p.n_mels + c] = v;
        }
        __syncthreads();
    }
}
