// Raw tcgen05.mma.kind::tf32 issue-rate probe: cycles per MMA (M=128, N=256, K=8) for K-major vs MN-major operands.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = (uint64_t)(layout & 7) << 61;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n, int amn, int bmn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__global__ void rate(int mn_major, int n, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 40960; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        uint64_t da, db;
        if (mn_major) { da = make_desc(smem_u32(smem), 4096, 512, 1); db = make_desc(smem_u32(smem) + 32768, 4096, 512, 1); }
        else          { da = make_desc(smem_u32(smem), 128, 1024, 0); db = make_desc(smem_u32(smem) + 32768, 128, 1024, 0); }
        uint32_t id = idesc(128, n, mn_major, mn_major);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(id), "r"(1u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840);
    for (int grid : {1, 148}) for (int n : {64, 128, 256}) for (int mn = 0; mn < 2; ++mn) {
        int iters = 2000;
        rate<<<grid, 128, 163840>>>(mn, n, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid=%3d N=%3d %s: %.1f cycles / MMA (M=128,K=8)\n", grid, n, mn ? "MN-major" : "K-major ", (double)mx / iters);
    }
    return 0;
}
