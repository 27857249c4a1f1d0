// wn_mel.cuh -- STFT -> mel -> dB -> normalise (the reference's utils/audio.py:69-75 melspectrogram chain:
// preemphasis :22-25, librosa.stft(n_fft, hop, win) :139-143, mel basis dot :181-199, _amp_to_db :201-203,
// _normalize :208-212) as one sm_100a kernel: one CTA per frame.
//
// librosa evaluates the FFT in float64 (the pre-emphasised signal is float64) and stores complex64; a float32
// FFT would miss the 1e-4 tolerance in bands far below the frame's peak (absolute error ~1e-7 * peak), so the
// FFT runs in fp64 in shared memory (B200 has full-rate fp64 units; 2048 points x 11 radix-2 stages is ~0.25
// MFLOP per frame), components are rounded to fp32 like complex64, and everything after is fp32 like numpy.
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
