// Bilateral-normal-filter regulariser as fused gather/reduce kernels.
//
// Replaces Loss.fn_bnf_detach_loss(pos, fn, mesh, ltype, loop) (util/loss.py:197-253), the `-CAD` term of the training
// step (sgcn.py:133-136, mgcn.py:144-147): ~40 ATen kernels per call (index gathers over f2f[F,3], exp, stack, sum,
// norm, five detached filter iterations) become
//   k_bnf_geom    face centroid + area from the (detached) positions              util/loss.py:203-205
//   k_bnf_dist    squared centroid distances to the three edge neighbours + the
//                 partial sums of their square roots (sigma_c)                     util/loss.py:212-216
//   k_bnf_iter    one filter iteration: gather 3 neighbour normals, weights
//                 exp(-dc/2 sigma_c^2) exp(-dn/2 sigma_s^2) area, normalise        util/loss.py:220-233   (x loop)
//   k_bnf_loss    distance of fn to the filtered normals, per-CTA partial sums    util/loss.py:235-249
//   k_bnf_final   fixed-order reduction -> loss                                    (one CTA)
// and one backward kernel (the filtered normals are detached: the gradient reaches `fn` through the final difference
// only).  All fp32 like the reference (pos / fn are the fp32 network output and its face normals); sums in fp64, rounded
// once.  Gather-bound: per iteration 12 B (f2f) + 12 B (centroid distances) + 4 x 12 B (normals) + 12 B (areas) per face.
//
// Reference quirks kept: an absent neighbour (f2f == -1) indexes the LAST face (Python's negative index) -- its centroid
// distance enters sigma_c, its area is zeroed (`no_neig`), so it never enters the filter.
#include "common.cuh"

namespace sgb {

constexpr int kBnfThreads = 256;
constexpr int kBnfRowsMax = 1024;

__device__ __forceinline__ double bnf_block_sum(double v, double* sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = kBnfThreads / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// fc = (p0 + p1 + p2) / 3,  fa = 0.5 sqrt(|cross(p1 - p0, p2 - p0)|^2 + 1e-12)
__global__ void __launch_bounds__(kBnfThreads) k_bnf_geom(const float* __restrict__ pos, int64_t ldp, const int64_t* __restrict__ faces, int64_t nf,
                                                          float* __restrict__ fc, float* __restrict__ fa) {
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        const float* a = pos + faces[f * 3] * ldp;
        const float* b = pos + faces[f * 3 + 1] * ldp;
        const float* c = pos + faces[f * 3 + 2] * ldp;
        float pa[3], e1[3], e2[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            pa[d] = __ldg(a + d);
            const float pb = __ldg(b + d), pc = __ldg(c + d);
            fc[f * 3 + d] = __fdiv_rn(__fadd_rn(__fadd_rn(pa[d], pb), pc), 3.0f);
            e1[d] = __fsub_rn(pb, pa[d]);
            e2[d] = __fsub_rn(pc, pa[d]);
        }
        const float cx = __fsub_rn(__fmul_rn(e1[1], e2[2]), __fmul_rn(e1[2], e2[1]));
        const float cy = __fsub_rn(__fmul_rn(e1[2], e2[0]), __fmul_rn(e1[0], e2[2]));
        const float cz = __fsub_rn(__fmul_rn(e1[0], e2[1]), __fmul_rn(e1[1], e2[0]));
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)), 1.0e-12f);
        fa[f] = __fmul_rn(0.5f, __fsqrt_rn(s));
    }
}

__device__ __forceinline__ int64_t bnf_wrap(int32_t nb, int64_t nf) { return nb < 0 ? nf + nb : nb; }     // python indexing of -1

__global__ void __launch_bounds__(kBnfThreads) k_bnf_dist(const float* __restrict__ fc, const int32_t* __restrict__ f2f, int64_t nf,
                                                          float* __restrict__ fcd, double* __restrict__ partials) {
    __shared__ double sh[kBnfThreads];
    double acc = 0.0;
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        const float c0 = fc[f * 3], c1 = fc[f * 3 + 1], c2 = fc[f * 3 + 2];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int64_t nb = bnf_wrap(__ldg(f2f + f * 3 + j), nf);
            const float d0 = __fsub_rn(fc[nb * 3], c0), d1 = __fsub_rn(fc[nb * 3 + 1], c1), d2 = __fsub_rn(fc[nb * 3 + 2], c2);
            const float dist = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
            fcd[f * 3 + j] = dist;
            acc += (double)__fsqrt_rn(__fadd_rn(dist, 1.0e-12f));
        }
    }
    acc = bnf_block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// scal[0] = sigma_c = sum sqrt(fc_dist + 1e-12) / (3 nf);  scal[1] = 2 sigma_c^2;  scal[2] = 2 sigma_s^2
__global__ void __launch_bounds__(kBnfThreads) k_bnf_sigma(const double* __restrict__ partials, int rows, int64_t nf, float* __restrict__ scal) {
    __shared__ double sh[kBnfThreads];
    double s = 0.0;
    for (int r = threadIdx.x; r < rows; r += kBnfThreads) s += partials[r];
    s = bnf_block_sum(s, sh);
    if (threadIdx.x == 0) {
        const float sigma_c = (float)(s / (double)(nf * 3));
        scal[0] = sigma_c;
        scal[1] = __fmul_rn(2.0f, __fmul_rn(sigma_c, sigma_c));
        scal[2] = 2.0f * (0.3f * 0.3f);               // 2 * sigma_s ** 2 evaluated like Python: 2 * 0.09
    }
}

// new_fn_f = normalise( sum_j wc_j ws_j fa_nb_j * cur_nb_j ),  wc = exp(-fc_dist / 2 sigma_c^2), ws = exp(-|cur_nb - cur_f|^2 / 2 sigma_s^2)
__global__ void __launch_bounds__(kBnfThreads) k_bnf_iter(const float* __restrict__ cur, const int32_t* __restrict__ f2f, const float* __restrict__ fcd,
                                                          const float* __restrict__ fa, const float* __restrict__ scal, int64_t nf,
                                                          float* __restrict__ nxt) {
    const float two_sc2 = __ldg(scal + 1);
    const float two_ss2 = (float)(2.0 * (0.3 * 0.3));
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        const float n0 = cur[f * 3], n1 = cur[f * 3 + 1], n2 = cur[f * 3 + 2];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int32_t nbr = __ldg(f2f + f * 3 + j);
            const int64_t nb = bnf_wrap(nbr, nf);
            const float m0 = cur[nb * 3], m1 = cur[nb * 3 + 1], m2 = cur[nb * 3 + 2];
            const float d0 = __fsub_rn(m0, n0), d1 = __fsub_rn(m1, n1), d2 = __fsub_rn(m2, n2);
            const float fn_dist = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
            const float wc = expf(__fdiv_rn(__fmul_rn(-1.0f, __ldg(fcd + f * 3 + j)), two_sc2));
            const float ws = expf(__fdiv_rn(__fmul_rn(-1.0f, fn_dist), two_ss2));
            const float nfa = nbr < 0 ? 0.0f : __ldg(fa + nb);                  // fa[f2f] * no_neig
            const float w = __fmul_rn(__fmul_rn(wc, ws), nfa);
            a0 = __fadd_rn(a0, __fmul_rn(w, m0));
            a1 = __fadd_rn(a1, __fmul_rn(w, m1));
            a2 = __fadd_rn(a2, __fmul_rn(w, m2));
        }
        const float len = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2)), 1.0e-12f)), 1.0e-12f);
        nxt[f * 3] = __fdiv_rn(a0, len);
        nxt[f * 3 + 1] = __fdiv_rn(a1, len);
        nxt[f * 3 + 2] = __fdiv_rn(a2, len);
    }
}

// ltype: 0 "mae", 1 "l1mae", 2 "rmse", 3 "l1rmse" (exactly as written in util/loss.py:235-249)
__global__ void __launch_bounds__(kBnfThreads) k_bnf_loss(const float* __restrict__ fn, const float* __restrict__ new_fn, int64_t nf, int ltype,
                                                          double* __restrict__ partials) {
    __shared__ double sh[kBnfThreads];
    double acc = 0.0;
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        const float d0 = __fsub_rn(new_fn[f * 3], fn[f * 3]), d1 = __fsub_rn(new_fn[f * 3 + 1], fn[f * 3 + 1]), d2 = __fsub_rn(new_fn[f * 3 + 2], fn[f * 3 + 2]);
        if (ltype == 0) {
            acc += (double)__fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), 1.0e-12f));
        } else if (ltype == 2) {
            acc += (double)__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
        } else {
            const float l1 = __fadd_rn(__fadd_rn(fabsf(d0), fabsf(d1)), fabsf(d2));
            acc += ltype == 1 ? (double)l1 : (double)__fmul_rn(l1, l1);
        }
    }
    acc = bnf_block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// out[0] = loss (rounded to fp32 like the reference's fp32 arithmetic), out[1] = the mean before the final sqrt
__global__ void __launch_bounds__(kBnfThreads) k_bnf_final(const double* __restrict__ partials, int rows, int64_t nf, int ltype, double* __restrict__ out) {
    __shared__ double sh[kBnfThreads];
    double s = 0.0;
    for (int r = threadIdx.x; r < rows; r += kBnfThreads) s += partials[r];
    s = bnf_block_sum(s, sh);
    if (threadIdx.x != 0) return;
    const float mean = (float)(s / (double)nf);
    float loss = mean;
    if (ltype == 2) loss = __fsqrt_rn(__fadd_rn(mean, 1.0e-12f));
    else if (ltype == 3) loss = __fsqrt_rn(__fadd_rn(__fmul_rn(mean, mean), 1.0e-12f));
    out[0] = (double)loss;
    out[1] = (double)mean;
}

// dloss/dfn (new_fn is detached):  mae -(d / sqrt(|d|^2 + 1e-12)) / F;  l1mae -sign(d) / F;  rmse -d / (F loss);
// l1rmse (mean / loss) * 2 l1 / F * -sign(d)           with d = new_fn - fn
__global__ void __launch_bounds__(kBnfThreads) k_bnf_bwd(const float* __restrict__ fn, const float* __restrict__ new_fn, int64_t nf, int ltype,
                                                         const double* __restrict__ out, const double* __restrict__ grad, float* __restrict__ dfn) {
    const double g = grad[0], inv_f = 1.0 / (double)nf, loss = out[0], mean = out[1];
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        float d[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) d[q] = __fsub_rn(new_fn[f * 3 + q], fn[f * 3 + q]);
        double k[3];
        if (ltype == 0) {
            const double r = sqrt((double)d[0] * d[0] + (double)d[1] * d[1] + (double)d[2] * d[2] + 1.0e-12);
#pragma unroll
            for (int q = 0; q < 3; ++q) k[q] = -(double)d[q] / r * inv_f;
        } else if (ltype == 2) {
#pragma unroll
            for (int q = 0; q < 3; ++q) k[q] = -(double)d[q] * inv_f / loss;
        } else {
            const double l1 = fabs((double)d[0]) + fabs((double)d[1]) + fabs((double)d[2]);
            const double c = ltype == 1 ? inv_f : (mean / loss) * 2.0 * l1 * inv_f;
#pragma unroll
            for (int q = 0; q < 3; ++q) k[q] = d[q] > 0.f ? -c : (d[q] < 0.f ? c : 0.0);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) dfn[f * 3 + q] = (float)(g * k[q]);
    }
}

static int bnf_grid(int64_t nf) {
    int64_t g = ceil_div(nf > 0 ? nf : 1, kBnfThreads);
    int64_t cap = (int64_t)num_sms() * 4;
    if (cap > kBnfRowsMax) cap = kBnfRowsMax;
    return (int)(g < cap ? g : cap);
}

}  // namespace sgb

using namespace sgb;

extern "C" size_t sgb_bnf_work_floats(int64_t nf) { return nf < 0 ? 0 : (size_t)nf * 13 + 8; }
extern "C" int sgb_bnf_partial_rows(void) { return kBnfRowsMax; }

extern "C" int sgb_bnf_loss_fwd(const float* pos, int64_t ldp, int64_t n, const int64_t* faces, const int32_t* f2f, int64_t nf, const float* fn,
                                int loop, int ltype, float* work, double* partials, float* new_fn, double* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(pos && ldp >= 3 && n > 0 && faces && f2f && nf > 0 && fn && work && partials && new_fn && out, "sgb_bnf_loss_fwd: bad argument");
    SGB_CHECK_ARG(loop >= 0 && ltype >= 0 && ltype <= 3, "sgb_bnf_loss_fwd: loop >= 0 and ltype in 0..3 (mae, l1mae, rmse, l1rmse)");
    float* scal = work;                 // [8]
    float* fc = work + 8;               // [3 nf]
    float* fa = fc + 3 * nf;            // [nf]
    float* fcd = fa + nf;               // [3 nf]
    float* ping = fcd + 3 * nf;         // [3 nf]
    float* pong = ping + 3 * nf;        // [3 nf]
    const int grid = bnf_grid(nf);
    k_bnf_geom<<<grid, kBnfThreads, 0, stream>>>(pos, ldp, faces, nf, fc, fa);
    SGB_CHECK_LAUNCH("k_bnf_geom");
    k_bnf_dist<<<grid, kBnfThreads, 0, stream>>>(fc, f2f, nf, fcd, partials);
    SGB_CHECK_LAUNCH("k_bnf_dist");
    k_bnf_sigma<<<1, kBnfThreads, 0, stream>>>(partials, grid, nf, scal);
    SGB_CHECK_LAUNCH("k_bnf_sigma");
    const float* cur = fn;
    for (int i = 0; i < loop; ++i) {
        float* nxt = (i == loop - 1) ? new_fn : ((i & 1) ? pong : ping);
        k_bnf_iter<<<grid, kBnfThreads, 0, stream>>>(cur, f2f, fcd, fa, scal, nf, nxt);
        SGB_CHECK_LAUNCH("k_bnf_iter");
        cur = nxt;
    }
    if (loop == 0) SGB_CUDA(cudaMemcpyAsync(new_fn, fn, (size_t)nf * 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    k_bnf_loss<<<grid, kBnfThreads, 0, stream>>>(fn, new_fn, nf, ltype, partials);
    SGB_CHECK_LAUNCH("k_bnf_loss");
    k_bnf_final<<<1, kBnfThreads, 0, stream>>>(partials, grid, nf, ltype, out);
    SGB_CHECK_LAUNCH("k_bnf_final");
    return SGB_OK;
}

extern "C" int sgb_bnf_loss_bwd(const float* fn, const float* new_fn, int64_t nf, int ltype, const double* out, const double* grad, float* dfn,
                                void* stream_) {
    SGB_CHECK_ARG(fn && new_fn && nf > 0 && out && grad && dfn && ltype >= 0 && ltype <= 3, "sgb_bnf_loss_bwd: bad argument");
    k_bnf_bwd<<<bnf_grid(nf), kBnfThreads, 0, (cudaStream_t)stream_>>>(fn, new_fn, nf, ltype, out, grad, dfn);
    SGB_CHECK_LAUNCH("k_bnf_bwd");
    return SGB_OK;
}
