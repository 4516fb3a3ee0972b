// Standalone probe: which shared-memory words does tcgen05.mma (kind::tf32, no swizzle) read for
// element (mn, k) of an MN-major A operand, as a function of the descriptor's LBO / SBO fields?
// A region is filled with word index values (exact in tf32 up to 2048); B is a K-major one-hot
// selecting one k; D[mn, 0] then reveals the word read for (mn, k).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n, int amn, int bmn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

// mode 0: A MN-major probe (B K-major one-hot).  mode 1: B MN-major probe (A K-major one-hot).
__global__ void probe(int mode, int kstar, uint32_t lbo, uint32_t sbo, int a_mn, int b_mn, uint32_t layout, float* out /*[128*16]*/) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    float* regA = reinterpret_cast<float*>(smem);              // 32 KB probe region
    float* regB = reinterpret_cast<float*>(smem + 32768);      // 32 KB region for the other operand
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += blockDim.x) { regA[i] = 0.f; regB[i] = 0.f; }
    __syncthreads();
    if (mode == 0) {
        for (int i = tid; i < 8192; i += blockDim.x) regA[i] = (float)(i % 2048);
        // B: K-major, N=16 rows, K=8: element (n, k) at (n/8)*SBOb + (k/4)*LBOb + (n%8)*16 + (k%4)*4, LBOb=128, SBOb=256
        if (tid == 0) regB[(kstar / 4) * 32 + (kstar % 4)] = 1.0f;   // n = 0
    } else {
        for (int i = tid; i < 8192; i += blockDim.x) regB[i] = (float)(i % 2048);
        // A: K-major, M=128, K=8: (m/8)*256 + (k/4)*128 + (m%8)*16 + (k%4)*4 bytes ; set A[m, kstar] = 1 for all m
        for (int m = tid; m < 128; m += blockDim.x) regA[((m / 8) * 256 + (kstar / 4) * 128 + (m % 8) * 16 + (kstar % 4) * 4) / 4] = 1.0f;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (tid == 0) {
        uint64_t da, db;
        if (mode == 0) { da = make_desc(smem_u32(regA), lbo, sbo, layout); db = make_desc(smem_u32(regB), 128, 256); }
        else           { da = make_desc(smem_u32(regA), 128, 256); db = make_desc(smem_u32(regB), lbo, sbo, layout); }
        uint32_t id = idesc(128, 16, a_mn, b_mn);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                     "l"(da), "l"(db), "r"(id), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid < 128) {
        const int warp = tid >> 5;
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[tid * 16 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, 128 * 16 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    std::vector<float> h(128 * 16);
    struct Cfg { int mode, amn, bmn; uint32_t lbo, sbo, layout; const char* name; };
    Cfg cfgs[] = {
        {0, 1, 0, 4096, 512, 1, "A MN-major SW128_BASE32B (LBO=4096,SBO=512)"},
        {0, 1, 0, 512, 4096, 1, "A MN-major SW128_BASE32B (LBO=512,SBO=4096)"},
        {1, 0, 1, 4096, 512, 1, "B MN-major SW128_BASE32B (LBO=4096,SBO=512)"},
    };
    for (auto& c : cfgs) {
        printf("=== %s\n", c.name);
        for (int kstar = 0; kstar < 8; kstar += 1) {
            cudaMemset(d_out, 0xFF, 128 * 16 * 4);
            probe<<<1, 128, 65536>>>(c.mode, kstar, c.lbo, c.sbo, c.amn, c.bmn, c.layout, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  k*=%d CUDA error: %s\n", kstar, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h.data(), d_out, 128 * 16 * 4, cudaMemcpyDeviceToHost);
            if (c.mode == 0) {
                printf("  k*=%d  word for mn=0,1,2,3,4,5,8,12,16,28,31,32,36,64,96,127:", kstar);
                int mns[] = {0, 1, 2, 3, 4, 5, 8, 12, 16, 28, 31, 32, 36, 64, 96, 127};
                for (int mn : mns) printf(" %g", h[mn * 16 + 0]);
                printf("\n");
            } else {
                printf("  k*=%d  word read for n=0..15 (row 0):", kstar);
                for (int n = 0; n < 16; ++n) printf(" %g", h[0 * 16 + n]);
                printf("\n");
            }
        }
    }
    return 0;
}
