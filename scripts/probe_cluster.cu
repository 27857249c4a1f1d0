// probe_cluster.cu -- r02 feasibility probe for the cluster / DSMEM sample chain (diagnostic, not product).
//   1. cudaOccupancyMaxActiveClusters for cluster sizes 2..16 at the kernel's shared-memory / thread budget
//   2. co-residency + SM / cluster placement of a 128-CTA grid of 16-CTA clusters
//   3. DSMEM one-to-one ping-pong (st.async + mbarrier complete_tx) round trip inside a cluster
//   4. "stage ring": 4 sender CTAs x 128 partial values -> 4 receiver CTAs (the layer-to-layer exchange of the
//      WaveNet chain), combine + named barrier + dense-like send, hop cost in cycles
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_cluster probe_cluster.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, long long budget = 400000000LL)
{
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > budget) return false;
    return true;
}
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar)
                 : "memory");
}

// ---- 2. placement / co-residency ---------------------------------------------------------------------
extern __shared__ __align__(128) unsigned char dsm[];
__global__ void place_kernel(int *counter, int *out, int total)
{
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) {
        out[blockIdx.x * 3 + 0] = (int)smid;
        out[blockIdx.x * 3 + 1] = (int)cluster_id_x();
        out[blockIdx.x * 3 + 2] = (int)cluster_ctarank();
        atomicAdd(counter, 1);
        long long t0 = clock64();
        while (atomicAdd(counter, 0) < total && clock64() - t0 < 2000000000LL) {}
        if (atomicAdd(counter, 0) < total) out[blockIdx.x * 3 + 0] = -1 - (int)smid;   // not co-resident
    }
    dsm[threadIdx.x] = 0;
}

// ---- 3. one-to-one DSMEM ping-pong: rank 0 <-> rank k --------------------------------------------------
__global__ void pingpong_kernel(int iters, long long *out)
{
    __shared__ uint64_t bar;
    __shared__ float slot;
    const uint32_t rank = cluster_ctarank();
    uint32_t csize;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0) {
        for (uint32_t k = 1; k < csize; ++k) {
            if (rank != 0 && rank != k) continue;
            const uint32_t peer = (rank == 0) ? k : 0;
            const uint32_t r_slot = mapa(smem_u32(&slot), peer), r_bar = mapa(smem_u32(&bar), peer);
            // phases of `bar` continue across partners on rank 0; rank k starts at phase 0
            static_cast<void>(0);
            long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                if (rank == 0) {
                    mbar_expect_tx(&bar, 4);
                    st_async_f32(r_slot, 1.0f, r_bar);
                    if (!mbar_wait(&bar, (uint32_t)((k - 1) * iters + i) & 1)) { out[64 + k] = -1; break; }
                } else {
                    mbar_expect_tx(&bar, 4);
                    if (!mbar_wait(&bar, (uint32_t)i & 1)) { out[64 + k] = -2; break; }
                    st_async_f32(r_slot, 1.0f, r_bar);
                }
            }
            if (rank == 0) out[k] = (clock64() - t0) / iters;
        }
    }
    cluster_sync_all();
}

// ---- 4. stage ring ------------------------------------------------------------------------------------
// cluster of CS CTAs = CS/4 stages x 4 CTAs.  Stage s receives 4 x 128 partials (one 512 B block per sender CTA),
// 128 threads sum them (+ dummy work of `work` dependent FMAs), named barrier, 128 "lead" threads post one value
// to each of the 4 CTAs of stage s+1 (wrap-around).  One token circulates; cycles per hop = total / (iters * stages).
template <int NT>
__global__ void ring_kernel(int iters, int work, long long *out)
{
    __shared__ uint64_t bar;
    __shared__ __align__(16) float inbox[4 * 128];
    __shared__ float xs[128];
    const uint32_t rank = cluster_ctarank();
    uint32_t csize;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const int stages = csize / 4, stage = rank / 4, m = rank % 4;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); mbar_expect_tx(&bar, 2048); }
    for (int i = tid; i < 512; i += NT) inbox[i] = 0.0f;
    __syncthreads();
    cluster_sync_all();
    const int nstage = (stage + 1) % stages;
    uint32_t r_in[4], r_bar[4];
    for (int j = 0; j < 4; ++j) {
        r_in[j] = mapa(smem_u32(&inbox[m * 128 + (tid & 127)]), nstage * 4 + j);
        r_bar[j] = mapa(smem_u32(&bar), nstage * 4 + j);
    }
    long long t0 = clock64();
    bool ok = true;
    for (int i = 0; i < iters && ok; ++i) {
        // stage 0 starts the token at i == 0 without waiting
        if (!(stage == 0 && i == 0)) {
            const uint32_t phase = (stage == 0) ? (uint32_t)(i - 1) & 1 : (uint32_t)i & 1;
            ok = mbar_wait(&bar, phase);
            if (tid == 0) mbar_expect_tx(&bar, 2048);
        }
        float v = 0.0f;
        if (tid < 128) {
            v = inbox[tid] + inbox[128 + tid] + inbox[256 + tid] + inbox[384 + tid];
            xs[tid] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        float a = xs[(tid * 7) & 127];
        for (int w = 0; w < work; ++w) a = fmaf(a, 1.0001f, 0.5f);
        if (tid < 128) {
#pragma unroll
            for (int j = 0; j < 4; ++j) st_async_f32(r_in[j], a * 0.25f, r_bar[j]);
        }
    }
    long long t1 = clock64();
    if (tid == 0 && rank == 0) { out[0] = ok ? (t1 - t0) / ((long long)iters * stages) : -1; }
    cluster_sync_all();
}

template <class K>
static cudaError_t launch_cluster(K kernel, int grid, int threads, int cluster, size_t smem, bool coop, void **args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = coop ? 2 : 1;
    return cudaLaunchKernelExC(&cfg, (const void *)kernel, args);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);

    // 1. occupancy
    CK(cudaFuncSetAttribute(place_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    for (int threads : {256, 512}) {
        for (int smem_kb : {64, 160, 200, 225}) {
            for (int cs : {1, 2, 4, 8, 16}) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs * 64);
                cfg.blockDim = dim3(threads);
                cfg.dynamicSmemBytes = (size_t)smem_kb * 1024;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int n = -1;
                cudaError_t e = cudaOccupancyMaxActiveClusters(&n, (const void *)place_kernel, &cfg);
                printf("occupancy threads=%d smem=%dKB cluster=%d -> max active clusters %d (%d CTAs) %s\n", threads, smem_kb, cs, n, n * cs,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
                if (e != cudaSuccess) cudaGetLastError();
            }
        }
    }

    // 2. placement
    int *d_counter, *d_out;
    CK(cudaMalloc(&d_counter, 4));
    CK(cudaMalloc(&d_out, 4096 * 4));
    for (int cs : {8, 16}) {
        for (int coop = 0; coop < 2; ++coop) {
            for (int grid : {128, 144}) {
                if (grid % cs) continue;
                CK(cudaMemset(d_counter, 0, 4));
                CK(cudaMemset(d_out, 0, 4096 * 4));
                int total = grid;
                void *args[] = {&d_counter, &d_out, &total};
                cudaError_t e = launch_cluster(place_kernel, grid, 512, cs, 200 * 1024, coop != 0, args);
                if (e == cudaSuccess) e = cudaDeviceSynchronize();
                printf("placement cluster=%d grid=%d coop=%d: %s\n", cs, grid, coop, cudaGetErrorString(e));
                if (e != cudaSuccess) { cudaGetLastError(); continue; }
                std::vector<int> h(grid * 3);
                CK(cudaMemcpy(h.data(), d_out, grid * 12, cudaMemcpyDeviceToHost));
                int bad = 0;
                for (int i = 0; i < grid; ++i) if (h[i * 3] < 0) ++bad;
                printf("  not co-resident: %d of %d\n", bad, grid);
                if (coop == 0) {
                    for (int c = 0; c < grid / cs; ++c) {
                        printf("  cluster %2d SMs:", c);
                        for (int i = 0; i < grid; ++i) if (h[i * 3 + 1] == c) printf(" %d", h[i * 3] < 0 ? -1 - h[i * 3] : h[i * 3]);
                        printf("\n");
                    }
                }
            }
        }
    }

    // 3. ping-pong
    long long *d_ll;
    CK(cudaMalloc(&d_ll, 256 * 8));
    CK(cudaFuncSetAttribute(pingpong_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int cs : {8, 16}) {
        CK(cudaMemset(d_ll, 0, 256 * 8));
        int iters = 2000;
        void *args[] = {&iters, &d_ll};
        cudaError_t e = launch_cluster(pingpong_kernel, cs, 32, cs, 0, false, args);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        printf("pingpong cluster=%d: %s\n", cs, cudaGetErrorString(e));
        if (e != cudaSuccess) { cudaGetLastError(); continue; }
        std::vector<long long> h(256);
        CK(cudaMemcpy(h.data(), d_ll, 256 * 8, cudaMemcpyDeviceToHost));
        printf("  DSMEM st.async+mbarrier round trip cycles rank0<->k:");
        for (int k = 1; k < cs; ++k) printf(" %lld%s", h[k], h[64 + k] ? "(!)" : "");
        printf("\n");
    }

    // 4. stage ring
    CK(cudaFuncSetAttribute(ring_kernel<256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(ring_kernel<512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int cs : {8, 16}) {
        for (int nt : {256, 512}) {
            for (int work : {0, 64, 256}) {
                CK(cudaMemset(d_ll, 0, 256 * 8));
                int iters = 4000;
                void *args[] = {&iters, &work, &d_ll};
                cudaError_t e = (nt == 256) ? launch_cluster(ring_kernel<256>, cs, 256, cs, 0, false, args)
                                            : launch_cluster(ring_kernel<512>, cs, 512, cs, 0, false, args);
                if (e == cudaSuccess) e = cudaDeviceSynchronize();
                long long h = 0;
                CK(cudaMemcpy(&h, d_ll, 8, cudaMemcpyDeviceToHost));
                printf("ring cluster=%d threads=%d dummy_fma=%d: %s, cycles per hop %lld (4x FMA latency x work = %d)\n", cs, nt, work,
                       cudaGetErrorString(e), h, work * 4);
                if (e != cudaSuccess) cudaGetLastError();
            }
        }
    }
    return 0;
}
