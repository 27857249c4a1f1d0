// probe_latency.cu -- r02 micro-latencies that bound one WaveNet layer on an SM (diagnostic, not product).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_latency probe_latency.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e_)); } while (0)

__device__ __forceinline__ float fdiv_rn(float a, float b) { return __fdiv_rn(a, b); }

// ---- dependent-chain latencies (1 warp) ----------------------------------------------------------------
__global__ void lat_kernel(long long *out, float seed)
{
    const int lane = threadIdx.x & 31;
    float a = seed + lane * 1e-3f;
    long long t0, t1;
    const int N = 256;
    // FFMA
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) a = __fmaf_rn(a, 1.000001f, 1e-7f);
    t1 = clock64(); if (lane == 0) out[0] = (t1 - t0);
    // FADD
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) a = __fadd_rn(a, 1e-7f);
    t1 = clock64(); if (lane == 0) out[1] = (t1 - t0);
    // SHFL.BFLY + FADD
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, 1 + (i & 15)));
    t1 = clock64(); if (lane == 0) out[2] = (t1 - t0);
    // SHFL only
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1 + (i & 15));
    t1 = clock64(); if (lane == 0) out[3] = (t1 - t0);
    // IEEE divide
    a = fabsf(a) + 1.5f;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) a = fdiv_rn(2.0f, __fadd_rn(a, 1.0f));
    t1 = clock64(); if (lane == 0) out[4] = (t1 - t0);
    // ex2.approx + rcp.approx
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float e, r;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(e, 1.0f)));
        a = r;
    }
    t1 = clock64(); if (lane == 0) out[5] = (t1 - t0);
    // select + shfl + fadd (one level of the transposing butterfly)
    float b = a * 0.5f;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const bool bit = (lane >> (i & 3)) & 1;
        const float keep = bit ? b : a, send = bit ? a : b;
        a = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 1 << (i & 3)));
        b = __fmul_rn(a, 0.999f);
    }
    t1 = clock64(); if (lane == 0) out[6] = (t1 - t0);
    out[32 + lane] = (long long)a + (long long)b;
}

// ---- shared-memory load patterns -------------------------------------------------------------------------
// every thread issues NL dependent-free LDS.128; distinct = number of distinct 16-byte addresses per warp
__global__ void lds_kernel(long long *out, int distinct, int nwarps_active)
{
    __shared__ __align__(16) float buf[4096];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 4096; i += blockDim.x) buf[i] = (float)i;
    __syncthreads();
    if (warp >= nwarps_active) return;
    const int NL = 32;
    const float4 *b4 = reinterpret_cast<const float4 *>(buf);
    const int base = (lane % distinct);
    float acc = 0.0f;
    long long t0 = clock64();
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
        float4 v[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) v[i] = b4[(base + i * distinct + rep * 7) & 1023];
#pragma unroll
        for (int i = 0; i < NL; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    long long t1 = clock64();
    if (lane == 0) out[warp] = (t1 - t0) / 8;
    if (acc == 1234.5f) out[63] = 1;
}

// LDS -> use latency (dependent pointer chase through shared memory)
__global__ void lds_lat_kernel(long long *out)
{
    __shared__ int nxt[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    int j = threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) j = nxt[j];
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    if (j == -1) out[1] = 1;
}

// ---- barrier cost: nthreads do bar.sync id in a loop ----------------------------------------------------------
__global__ void bar_kernel(long long *out, int mode)
{
    const int tid = threadIdx.x;
    long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < 256; ++i) asm volatile("bar.sync 1, 256;" ::: "memory");
    } else if (mode == 1) {
        for (int i = 0; i < 256; ++i) {
            int r;
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbarrier.red.or.pred q, 1, 256, p;\n\tselp.s32 %0, 1, 0, q;\n\t}" : "=r"(r) : "r"(0) : "memory");
            if (r) break;
        }
    } else {
        __shared__ float s[256];
        float a = (float)tid;
        for (int i = 0; i < 256; ++i) {
            s[tid] = a;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            a = s[(tid + 33) & 255] + 1.0f;
        }
        if (a == 3.0f) out[5] = 1;
    }
    long long t1 = clock64();
    if (tid == 0) out[mode] = (t1 - t0);
}

int main()
{
    long long *d, h[64];
    CK(cudaMalloc(&d, 64 * 8));
    CK(cudaMemset(d, 0, 64 * 8));
    lat_kernel<<<1, 32>>>(d, 0.37f);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost));
    const char *nm[] = {"FFMA dependent", "FADD dependent", "SHFL.BFLY+FADD", "SHFL.BFLY", "IEEE fdiv(2, a+1)", "ex2.approx + fadd + rcp.approx", "select+shfl+fadd+fmul (butterfly level)"};
    for (int i = 0; i < 7; ++i) printf("latency %-42s %.1f cycles\n", nm[i], h[i] / 256.0);

    lds_lat_kernel<<<1, 64>>>(d);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost));
    printf("latency %-42s %.1f cycles\n", "LDS.32 pointer chase", h[0] / 256.0);

    for (int nw : {1, 4, 8}) {
        for (int distinct : {1, 2, 4, 8, 16, 32}) {
            CK(cudaMemset(d, 0, 64 * 8));
            lds_kernel<<<1, 256>>>(d, distinct, nw);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            printf("LDS.128 x32 per thread, %d warps, %2d distinct 16B addresses per warp: %lld cycles (%.1f per LDS per warp, %.2f wavefront-cycles per LDS)\n",
                   nw, distinct, mx, mx / 32.0, mx / 32.0 / nw);
        }
    }
    for (int mode = 0; mode < 3; ++mode) {
        CK(cudaMemset(d, 0, 64 * 8));
        bar_kernel<<<1, 256>>>(d, mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost));
        printf("barrier mode %d (0 bar.sync 256, 1 barrier.red.or, 2 STS+bar.sync+LDS): %.1f cycles per iteration\n", mode, h[mode] / 256.0);
    }
    return 0;
}
