// Input preparation of the SGCN / MGCN forward as two kernels.
//
// Replaces, per forward (util/networks.py:67-79, util/meshnet.py:282-293):
//   z_min, z_max = min / max over the vertices of z1            (2 reductions)
//   z_sc = max(z_max - z_min);  zc = (z_min + z_max) * 0.5       (3 small ops)
//   z1 = (z1 - zc) / z_sc;  z1 = dm * z1;  x = cat([z1, dm], 1)  (4 elementwise ops + a concatenation)
// -- ~10 ATen launches on [N,3] / [N,4] tensors, which is what a launch-bound small-mesh step is made of -- by
//   k_bbox       per-CTA min / max -> six atomicMin / atomicMax on an order-preserving integer encoding of the floats
//                (min / max are order-independent: deterministic)
//   k_input_prep the bounding-box scalars + the normalise / mask / concatenate pass, one float4 store per vertex.
// Same arithmetic, same rounding (separately rounded sub / div / mul): bit-identical to the torch ops.  HBM-bound:
// reads 12 B (+4 B mask) twice, writes 16 B per vertex.
#include "common.cuh"

namespace sgb {

// order-preserving map float -> uint32 (negative floats reversed), so that unsigned atomicMin / atomicMax order like floats
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// box[0..2] = ord(min), box[3..5] = ord(max); initialised by k_bbox_init
__global__ void k_bbox_init(uint32_t* __restrict__ box) {
    if (threadIdx.x < 3) box[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) box[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) k_bbox(const float* __restrict__ z, int64_t ldz, int64_t n, uint32_t* __restrict__ box) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float x = __ldg(z + v * ldz + d);
            lo[d] = fminf(lo[d], x);
            hi[d] = fmaxf(hi[d], x);
        }
    }
    __shared__ uint32_t s[6];
    if (threadIdx.x < 3) s[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) s[threadIdx.x] = 0u;
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const uint32_t a = __reduce_min_sync(0xffffffffu, f2ord(lo[d]));
        const uint32_t b = __reduce_max_sync(0xffffffffu, f2ord(hi[d]));
        if ((threadIdx.x & 31) == 0) { atomicMin(&s[d], a); atomicMax(&s[3 + d], b); }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&box[threadIdx.x], s[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&box[threadIdx.x], s[threadIdx.x]);
}

// x[v] = (dm * ((z - zc) / z_sc), dm);  zc = (z_min + z_max) * 0.5, z_sc = max_d (z_max - z_min)
__global__ void __launch_bounds__(256) k_input_prep(const float* __restrict__ z, int64_t ldz, int64_t n, const float* __restrict__ dm,
                                                    const uint32_t* __restrict__ box, float* __restrict__ x, float* __restrict__ stats_out) {
    float zc[3], sc = -INFINITY;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float lo = ord2f(box[d]), hi = ord2f(box[3 + d]);
        zc[d] = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
        sc = fmaxf(sc, __fsub_rn(hi, lo));
    }
    if (stats_out && blockIdx.x == 0 && threadIdx.x == 0) {
        stats_out[0] = zc[0]; stats_out[1] = zc[1]; stats_out[2] = zc[2]; stats_out[3] = sc;
    }
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        const float m = dm ? __ldg(dm + v) : 1.0f;
        float4 o;
        o.x = __fmul_rn(m, __fdiv_rn(__fsub_rn(__ldg(z + v * ldz + 0), zc[0]), sc));
        o.y = __fmul_rn(m, __fdiv_rn(__fsub_rn(__ldg(z + v * ldz + 1), zc[1]), sc));
        o.z = __fmul_rn(m, __fdiv_rn(__fsub_rn(__ldg(z + v * ldz + 2), zc[2]), sc));
        o.w = m;
        st4(x + v * 4, o);
    }
}

}  // namespace sgb

using namespace sgb;

extern "C" int sgb_input_prep(const float* z1, int64_t ldz, int64_t n, const float* dm, float* x, float* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(z1 && x && scratch && n > 0 && ldz >= 3, "sgb_input_prep: bad argument");
    SGB_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 3) == 0, "sgb_input_prep: x must be 16-byte aligned");
    uint32_t* box = reinterpret_cast<uint32_t*>(scratch);
    const int grid = (int)min64(ceil_div(n, 256), (int64_t)num_sms() * 8);
    k_bbox_init<<<1, 32, 0, stream>>>(box);
    SGB_CHECK_LAUNCH("k_bbox_init");
    k_bbox<<<grid, 256, 0, stream>>>(z1, ldz, n, box);
    SGB_CHECK_LAUNCH("k_bbox");
    k_input_prep<<<grid, 256, 0, stream>>>(z1, ldz, n, dm, box, x, scratch + 8);
    SGB_CHECK_LAUNCH("k_input_prep");
    return SGB_OK;
}
