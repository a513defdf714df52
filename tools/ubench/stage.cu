// How fast can one CTA pull a 224 KB table from L2 into shared memory?  (sm_100a)
// Methods: 0 = cp.async.bulk (TMA bulk, 32 KB chunks, one issuing thread), 1 = bulk with 8 issuing
// threads / 4 KB chunks, 2 = 1024 threads x ld.global.nc.v4 + st.shared.v4, 3 = cp.async 16 B (LDGSTS).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int METHOD>
__global__ void __launch_bounds__(1024, 1) stage(const uint8_t* __restrict__ src, uint32_t bytes, long long* cyc, uint32_t* sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    long long t0 = clock64();
    if (METHOD == 0 || METHOD == 1) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const uint32_t chunk = METHOD == 0 ? 32768u : 4096u;
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        __syncthreads();
        const uint32_t issuers = METHOD == 0 ? 1u : 32u;
        if (threadIdx.x < issuers)
            for (uint32_t off = threadIdx.x * chunk; off < bytes; off += issuers * chunk)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                             "r"(s32(smem + off)), "l"(src + off), "r"(chunk), "r"(s32(&bar)) : "memory");
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(&bar)) : "memory");
    } else if (METHOD == 2) {
        const uint4* s = (const uint4*)src; uint4* d = (uint4*)smem;
        for (uint32_t i = threadIdx.x; i < bytes / 16; i += 1024 * 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) if (i + u * 1024 < bytes / 16) v[u] = __ldg(s + i + u * 1024);
#pragma unroll
            for (int u = 0; u < 4; u++) if (i + u * 1024 < bytes / 16) d[i + u * 1024] = v[u];
        }
        __syncthreads();
    } else {
        for (uint32_t i = threadIdx.x; i < bytes / 16; i += 1024)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(smem + i * 16)), "l"(src + i * 16) : "memory");
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (threadIdx.x == 5) sink[blockIdx.x] = ((uint32_t*)smem)[(bytes / 4 - 1) & 0xffff];
}

template <int M> void run(const uint8_t* src, uint32_t bytes, int grid, long long* cyc, uint32_t* sink, const char* name)
{
    cudaFuncSetAttribute(stage<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 229376);
    for (int rep = 0; rep < 3; rep++) stage<M><<<grid, 1024, 229376>>>(src, bytes, cyc, sink);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int rep = 0; rep < 20; rep++) stage<M><<<grid, 1024, 229376>>>(src, bytes, cyc, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[256]; cudaMemcpy(h, cyc, 8 * grid, cudaMemcpyDeviceToHost);
    double avg = 0, mx = 0; for (int i = 0; i < grid; i++) { avg += h[i]; if (h[i] > mx) mx = h[i]; } avg /= grid;
    cudaError_t e = cudaGetLastError();
    printf("%-28s bytes %6u grid %3d: %8.0f cycles avg %8.0f max (%.1f B/clk/SM)  %.2f us per launch  %s\n", name, bytes, grid, avg, mx,
           bytes / avg, ms * 1000 / 20, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    uint8_t* src; long long* cyc; uint32_t* sink;
    cudaMalloc(&src, 1 << 20); cudaMemset(src, 1, 1 << 20); cudaMalloc(&cyc, 8 * 256); cudaMalloc(&sink, 4 * 256);
    for (int grid : {1, 32, 148}) for (uint32_t bytes : {229376u, 131072u, 32768u}) {
        run<0>(src, bytes, grid, cyc, sink, "bulk 32KB x1 thread");
        run<1>(src, bytes, grid, cyc, sink, "bulk 4KB x32 threads");
        run<2>(src, bytes, grid, cyc, sink, "ldg.v4 + sts.v4 x1024");
        run<3>(src, bytes, grid, cyc, sink, "cp.async 16B x1024");
    }
    return 0;
}
