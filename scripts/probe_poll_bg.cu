// probe_poll_bg.cu -- r02: one warp's polling round (16 x ld.relaxed.gpu.u64 per lane, data present in L2) while other warps spin.
// Backgrounds: (a) none, (b) other SMs' CTAs polling never-arriving L2 words, (c) same-SM warps spinning on mbarrier.try_wait,
// (d) same-SM warps polling L2 words, (e) b+c.  Explains the 3.5k-cycle sampler poll seen in the live kernel (probe_poll: 760 alone).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_poll_bg probe_poll_bg.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e_)); } while (0)
typedef unsigned long long u64;
__device__ __forceinline__ u64 ld_relaxed(const u64 *p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_flag(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void fill_kernel(u64 *buf, int n) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = ((u64)7 << 32) | (unsigned)i; }

// block 0: warp 0 measures; warps 1.. are the same-SM background (smode 0 idle at a barrier, 1 mbarrier.try_wait spin, 2 L2 poll, 3 nanosleep poll)
// blocks 1..: other-SM background (gmode 0 exit at once, 1 poll L2 words with `gl` loads in flight, 2 same with nanosleep back-off)
__global__ void probe(const u64 *buf, u64 *never, unsigned *stop, long long *out, int smode, int gmode, int gl, int slot)
{
    __shared__ uint64_t bar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    __syncthreads();
    if (blockIdx.x == 0 && warp == 0) {
        long long best = 1ll << 60, sum = 0, first = 0;
        u64 chk = 0;
        for (int rep = 0; rep < 36; ++rep) {
            const u64 *base = buf + (size_t)(rep & 7) * 4096 + lane;
            __syncwarp();
            long long t0 = clock64(), tf = 0;
            u64 v[16];
            if (lane < 30) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = ld_relaxed(base + (size_t)i * 30);
                chk += v[0];
                tf = clock64();
#pragma unroll
                for (int i = 1; i < 16; ++i) chk += v[i];
            }
            __syncwarp();
            long long t1 = clock64();
            if (rep >= 4) { sum += t1 - t0; first += tf - t0; best = (t1 - t0 < best) ? t1 - t0 : best; }
            for (int k = 0; k < 50; ++k) asm volatile("nanosleep.u32 20;");
        }
        if (lane == 0) { out[slot * 4] = best; out[slot * 4 + 1] = sum / 32; out[slot * 4 + 2] = first / 32; *(volatile unsigned *)stop = 1u; asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory"); }
        if (chk == 1) out[63] = 1;
        return;
    }
    if (blockIdx.x == 0) {
        if (smode == 0) return;
        if (smode == 1) {
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            return;
        }
        const u64 *w = never + 65536 + (size_t)threadIdx.x;
        while (ld_flag(stop) == 0) {
            u64 a = ld_relaxed(w), b = ld_relaxed(w + 1024), c = ld_relaxed(w + 2048), d = ld_relaxed(w + 3072);
            if ((a | b | c | d) == 99) break;
            if (smode == 3) asm volatile("nanosleep.u32 100;");
        }
        return;
    }
    if (gmode == 0) return;
    const u64 *w = never + (size_t)blockIdx.x * 256 + threadIdx.x;
    unsigned it = 0;
    while (true) {
        u64 a = 0;
        for (int i = 0; i < gl; ++i) a |= ld_relaxed(w + (size_t)i * 40000);
        if (a == 99) break;
        if ((++it & 15) == 0 && ld_flag(stop)) break;
        if (gmode == 2) asm volatile("nanosleep.u32 100;");
    }
}

int main()
{
    u64 *buf, *never;
    unsigned *stop;
    long long *d, h[4];
    const int n = 9 * 4096;
    CK(cudaMalloc(&buf, n * 8));
    CK(cudaMalloc(&never, 400000 * 8));
    CK(cudaMemset(never, 0, 400000 * 8));
    CK(cudaMalloc(&stop, 4));
    CK(cudaMalloc(&d, 64 * 8));
    fill_kernel<<<32, 256>>>(buf, n);
    CK(cudaDeviceSynchronize());
    struct Cfg { int smode, gmode, gl, blocks, threads; const char *name; } cfgs[] = {
        {0, 0, 0, 1, 384, "alone"},
        {1, 0, 0, 1, 384, "same SM: 11 warps in mbarrier.try_wait"},
        {2, 0, 0, 1, 384, "same SM: 11 warps polling L2 (4 loads in flight)"},
        {3, 0, 0, 1, 384, "same SM: 11 warps polling L2 with nanosleep 100"},
        {0, 1, 1, 137, 128, "136 other CTAs x 128 threads polling L2, 1 load in flight"},
        {0, 1, 4, 137, 128, "136 other CTAs x 128 threads polling L2, 4 loads in flight"},
        {0, 1, 4, 137, 384, "136 other CTAs x 384 threads polling L2, 4 loads in flight"},
        {0, 2, 4, 137, 384, "136 other CTAs x 384 threads polling L2, 4 loads, nanosleep 100"},
        {0, 1, 4, 137, 32, "136 other CTAs x 32 threads polling L2, 4 loads in flight"},
        {1, 1, 4, 137, 384, "both: try_wait on the SM + 136 x 384 L2 pollers"},
    };
    for (auto &c : cfgs) {
        CK(cudaMemset(stop, 0, 4));
        probe<<<c.blocks, c.threads>>>(buf, never, stop, d, c.smode, c.gmode, c.gl, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost));
        printf("%-70s round of 16 loads/lane: best %5lld mean %5lld, first word after %5lld cycles\n", c.name, h[0], h[1], h[2]);
    }
    return 0;
}
