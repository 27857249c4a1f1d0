// probe_poll.cu -- r02: cost of one polling round over flagged 8-byte words that are already in L2 (diagnostic, not product).
// Question: why does the layer-0 sampler warp need ~3.5k cycles for 16 x ld.relaxed.gpu.u64 per lane with the data present?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_poll probe_poll.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e_)); } while (0)
typedef unsigned long long u64;

__device__ __forceinline__ u64 ld_relaxed(const u64 *p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ u64 ld_volatile(const u64 *p) { u64 v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ u64 ld_cg(const u64 *p) { u64 v; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ u64 ld_acquire(const u64 *p) { u64 v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void ld_relaxed_v2(const u64 *p, u64 &a, u64 &b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }

__global__ void fill_kernel(u64 *buf, int n) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = ((u64)7 << 32) | (unsigned)i; }

// mode 0 relaxed.gpu, 1 volatile, 2 cg, 3 acquire.gpu ; K loads per lane in flight, `lanes` active lanes, stride between a lane's words = O words
template <int K, int MODE>
__global__ void poll_kernel(const u64 *buf, long long *out, int lanes, int O, int nwarps, int slot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= nwarps) return;
    long long best = 1ll << 60, sum = 0;
    u64 chk = 0;
    for (int rep = 0; rep < 20; ++rep) {
        const u64 *base = buf + (size_t)(rep & 7) * 4096 + warp * K * O + lane;
        __syncwarp();
        long long t0 = clock64();
        u64 v[K];
        if (lane < lanes) {
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const u64 *a = base + (size_t)i * O;
                v[i] = MODE == 0 ? ld_relaxed(a) : MODE == 1 ? ld_volatile(a) : MODE == 2 ? ld_cg(a) : ld_acquire(a);
            }
#pragma unroll
            for (int i = 0; i < K; ++i) chk += v[i];
        }
        __syncwarp();
        long long t1 = clock64();
        if (rep >= 4) { sum += t1 - t0; best = (t1 - t0 < best) ? t1 - t0 : best; }
    }
    if (lane == 0 && warp == 0) { out[slot * 2] = best; out[slot * 2 + 1] = sum / 16; }
    if (chk == 1) out[63] = 1;
}

template <int K>
__global__ void poll_v2_kernel(const u64 *buf, long long *out, int lanes, int slot)
{
    const int lane = threadIdx.x & 31;
    long long best = 1ll << 60, sum = 0;
    u64 chk = 0;
    for (int rep = 0; rep < 20; ++rep) {
        const u64 *base = buf + (size_t)(rep & 7) * 4096 + lane * 2;
        __syncwarp();
        long long t0 = clock64();
        if (lane < lanes) {
            u64 a[K], b[K];
#pragma unroll
            for (int i = 0; i < K; ++i) ld_relaxed_v2(base + (size_t)i * 64, a[i], b[i]);
#pragma unroll
            for (int i = 0; i < K; ++i) chk += a[i] + b[i];
        }
        __syncwarp();
        long long t1 = clock64();
        if (rep >= 4) { sum += t1 - t0; best = (t1 - t0 < best) ? t1 - t0 : best; }
    }
    if (lane == 0) { out[slot * 2] = best; out[slot * 2 + 1] = sum / 16; }
    if (chk == 1) out[63] = 1;
}

template <int K, int MODE>
static void run(const u64 *buf, long long *d, int lanes, int O, int nwarps, const char *name)
{
    long long h[2];
    poll_kernel<K, MODE><<<1, 32 * nwarps>>>(buf, d, lanes, O, nwarps, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("%-14s K=%2d loads/lane, %2d lanes, stride %2d words, %d warp(s): best %5lld  mean %5lld cycles per round\n", name, K, lanes, O, nwarps, h[0], h[1]);
}

int main()
{
    u64 *buf;
    long long *d;
    const int n = 8 * 4096 + 4096;
    CK(cudaMalloc(&buf, n * 8));
    CK(cudaMalloc(&d, 64 * 8));
    fill_kernel<<<32, 256>>>(buf, n);
    CK(cudaDeviceSynchronize());
    run<1, 0>(buf, d, 30, 30, 1, "relaxed.gpu");
    run<2, 0>(buf, d, 30, 30, 1, "relaxed.gpu");
    run<4, 0>(buf, d, 30, 30, 1, "relaxed.gpu");
    run<8, 0>(buf, d, 30, 30, 1, "relaxed.gpu");
    run<16, 0>(buf, d, 30, 30, 1, "relaxed.gpu");
    run<16, 0>(buf, d, 30, 32, 1, "relaxed.gpu");
    run<16, 0>(buf, d, 32, 32, 1, "relaxed.gpu");
    run<16, 0>(buf, d, 16, 16, 1, "relaxed.gpu");
    run<16, 0>(buf, d, 4, 4, 1, "relaxed.gpu");
    run<16, 0>(buf, d, 1, 1, 1, "relaxed.gpu");
    run<4, 0>(buf, d, 30, 30, 4, "relaxed.gpu");
    run<8, 0>(buf, d, 30, 30, 2, "relaxed.gpu");
    run<16, 1>(buf, d, 30, 30, 1, "volatile");
    run<16, 2>(buf, d, 30, 30, 1, "ld.cg");
    run<4, 2>(buf, d, 30, 30, 1, "ld.cg");
    run<16, 3>(buf, d, 30, 30, 1, "acquire.gpu");
    run<4, 3>(buf, d, 30, 30, 1, "acquire.gpu");
    long long h[2];
    poll_v2_kernel<8><<<1, 32>>>(buf, d, 30, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("relaxed.gpu.v2.u64 K=8 x 16B per lane, 30 lanes: best %lld mean %lld\n", h[0], h[1]);
    poll_v2_kernel<4><<<1, 32>>>(buf, d, 32, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("relaxed.gpu.v2.u64 K=4 x 16B per lane, 32 lanes: best %lld mean %lld\n", h[0], h[1]);
    return 0;
}
