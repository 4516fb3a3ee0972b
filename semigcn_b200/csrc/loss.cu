// Step-path mesh losses as fused gather/reduce kernels.
//
// Replaces, per training step (sgcn.py:130-137, mgcn.py:137-143):
//   Models.compute_fn(pos, faces)                 util/models.py:121-126   gather 3 vertices/face -> cross -> normalise
//   Loss.mask_pos_rec_loss(pos, ini_vs, v_mask)   util/loss.py:14-34       sqrt(mean_{mask} |real - pred|^2 + 1e-6)
//   Loss.mask_norm_rec_loss(norm, fn, f_mask)     util/loss.py:78-107      mean_{mask} sum_d |pred - real|
//   Loss.mesh_laplacian_loss(pos, mesh)           util/loss.py:60-76       sqrt(mean |pos - Adj pos / deg|^2 + 1e-12)
// (~10 ATen kernels + boolean-mask indexing with a host sync each) by one forward pass, one tiny finalisation
// and one backward pass.  Arithmetic types follow the reference: pos and the face normal are fp32
// (compute_fn runs on the fp32 network output), the differences against the targets and all sums are
// fp64 when the targets are fp64 (sgcn.py feeds float64 numpy arrays, sgcn.py:127,132) -- sums are always
// accumulated in fp64 here, then rounded to the target type.
//
// Deterministic: per-CTA partial sums reduced in a fixed order; the backward pass is a per-vertex gather
// over the vertex->face incidence CSR (sgb_incidence_build), no floating-point atomics.
// HBM-bound: forward reads pos 12 B/vertex (+ gathered 36 B/face, L2-resident), targets 24|12 B per
// vertex and per face, masks 1 B; backward the same plus the incidence (4 B per corner), writes 12 B/vertex.
#include "common.cuh"

namespace sgb {

int exclusive_scan_i32(const int32_t* cnt, int64_t n, int32_t* tsum, int32_t* rowptr, cudaStream_t stream);   // graph_build.cu

constexpr int kLossThreads = 256;
constexpr int kLossRowsMax = 1024;

struct LossArgs {
    const float* pos; int64_t ldp; int64_t n;
    const void* tpos; int t64;                 // target positions [n,3], fp64 (t64 = 1) or fp32
    const uint8_t* vmask;                      // [n] or NULL (all)
    const int64_t* faces; int64_t nf;          // [nf,3]
    const void* tfn;                           // target face normals [nf,3], same type as tpos
    const uint8_t* fmask;                      // [nf] or NULL
};

__device__ __forceinline__ double tget(const void* p, int t64, int64_t i) {
    return t64 ? reinterpret_cast<const double*>(p)[i] : (double)reinterpret_cast<const float*>(p)[i];
}

// unit normal of face (a, b, c) exactly as compute_fn: cross(b - a, c - a) / sqrt(sum(n^2)), fp32, no FMA contraction
__device__ __forceinline__ void face_normal(const float* __restrict__ pos, int64_t ldp, int64_t ia, int64_t ib, int64_t ic, float n[3],
                                            float e1[3], float e2[3], float& len) {
    const float* a = pos + ia * ldp;
    const float* b = pos + ib * ldp;
    const float* c = pos + ic * ldp;
#pragma unroll
    for (int d = 0; d < 3; ++d) { e1[d] = __fsub_rn(__ldg(b + d), __ldg(a + d)); e2[d] = __fsub_rn(__ldg(c + d), __ldg(a + d)); }
    float cx = __fsub_rn(__fmul_rn(e1[1], e2[2]), __fmul_rn(e1[2], e2[1]));
    float cy = __fsub_rn(__fmul_rn(e1[2], e2[0]), __fmul_rn(e1[0], e2[2]));
    float cz = __fsub_rn(__fmul_rn(e1[0], e2[1]), __fmul_rn(e1[1], e2[0]));
    len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
    n[0] = __fdiv_rn(cx, len); n[1] = __fdiv_rn(cy, len); n[2] = __fdiv_rn(cz, len);
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
    // fixed-order tree over the CTA
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = kLossThreads / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    double r = sh[0];
    __syncthreads();
    return r;
}

// partials[gridDim.x][4] = (sum sq pos err, count_v, sum L1 normal err, count_f); optionally writes the normals
__global__ void __launch_bounds__(kLossThreads) k_step_loss_fwd(const LossArgs a, float* __restrict__ fn_out, double* __restrict__ partials) {
    __shared__ double sh[kLossThreads];
    double sq = 0.0, cv = 0.0, l1 = 0.0, cf = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (a.tpos) {
        for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < a.n; v += stride) {
            if (a.vmask && !a.vmask[v]) continue;
            cv += 1.0;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float p = __ldg(a.pos + v * a.ldp + d);
                if (a.t64) { const double df = tget(a.tpos, 1, v * 3 + d) - (double)p; sq += df * df; }
                else { const float df = fabsf(__fsub_rn(reinterpret_cast<const float*>(a.tpos)[v * 3 + d], p)); sq += (double)__fmul_rn(df, df); }
            }
        }
    }
    if (a.faces) {
        for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < a.nf; f += stride) {
            const bool on = !a.fmask || a.fmask[f];
            if (!on && !fn_out) continue;
            float n[3], e1[3], e2[3], len;
            face_normal(a.pos, a.ldp, a.faces[f * 3], a.faces[f * 3 + 1], a.faces[f * 3 + 2], n, e1, e2, len);
            if (fn_out) { fn_out[f * 3] = n[0]; fn_out[f * 3 + 1] = n[1]; fn_out[f * 3 + 2] = n[2]; }
            if (on && a.tfn) {
                cf += 1.0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (a.t64) l1 += fabs((double)n[d] - tget(a.tfn, 1, f * 3 + d));
                    else l1 += (double)fabsf(__fsub_rn(n[d], reinterpret_cast<const float*>(a.tfn)[f * 3 + d]));
                }
            }
        }
    }
    sq = block_sum(sq, sh); cv = block_sum(cv, sh); l1 = block_sum(l1, sh); cf = block_sum(cf, sh);
    if (threadIdx.x == 0) {
        double* o = partials + (int64_t)blockIdx.x * 4;
        o[0] = sq; o[1] = cv; o[2] = l1; o[3] = cf;
    }
}

// out[0] = loss_p, out[1] = loss_n, out[2] = count_v, out[3] = count_f
__global__ void __launch_bounds__(kLossThreads) k_step_loss_finalize(const double* __restrict__ partials, int rows, int t64, double* __restrict__ out) {
    // one CTA: thread r sums rows r, r + 256, ... (fixed order), then the fixed-order tree of block_sum -> deterministic
    __shared__ double sh[kLossThreads];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = threadIdx.x; r < rows; r += kLossThreads)
        for (int j = 0; j < 4; ++j) s[j] += partials[(int64_t)r * 4 + j];
    for (int j = 0; j < 4; ++j) s[j] = block_sum(s[j], sh);
    if (threadIdx.x != 0) return;
    double lp = sqrt(s[0] / s[1] + 1.0e-6);            // 0/0 -> NaN like the reference's empty mask
    double ln = s[2] / s[3];
    if (!t64) { lp = (double)sqrtf((float)(s[0] / s[1]) + 1.0e-6f); ln = (double)(float)ln; }
    out[0] = lp; out[1] = ln; out[2] = s[1]; out[3] = s[3];
}

struct LossBwdArgs {
    LossArgs a;
    const int32_t* inc_rowptr; const int32_t* inc;     // vertex -> (3*face + corner)
    const double* out;                                 // forward results
    const double* grads;                               // [2] upstream d/d loss_p, d/d loss_n
    float* dpos; int64_t lddpos;
};

// One thread per vertex: position term + the normal-loss contributions of every incident face corner.
//   n = c/|c|, c = e1 x e2:   dc = (g - n (n.g)) / |c|,  de1 = e2 x dc,  de2 = dc x e1,
//   corner 0: -(de1 + de2), corner 1: de1, corner 2: de2;   g = sign(n - target) * mask / count_f
__global__ void __launch_bounds__(kLossThreads) k_step_loss_bwd(const LossBwdArgs b) {
    const LossArgs& a = b.a;
    const double gp = b.grads[0], gn = b.grads[1];
    const double lp = b.out[0], cv = b.out[2], cf = b.out[3];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < a.n; v += stride) {
        double g[3] = {0.0, 0.0, 0.0};
        if (a.tpos && (!a.vmask || a.vmask[v])) {
            // d sqrt(S/c + eps) / d pos = (pos - target) / (c * loss_p)
            const double k = gp / (cv * lp);
#pragma unroll
            for (int d = 0; d < 3; ++d) g[d] = k * ((double)__ldg(a.pos + v * a.ldp + d) - tget(a.tpos, a.t64, v * 3 + d));
        }
        if (a.faces && a.tfn) {
            const double kf = gn / cf;
            for (int p = b.inc_rowptr[v]; p < b.inc_rowptr[v + 1]; ++p) {
                const int fc = b.inc[p];
                const int64_t f = fc / 3;
                const int corner = fc - (int)f * 3;
                if (a.fmask && !a.fmask[f]) continue;
                float n[3], e1[3], e2[3], len;
                face_normal(a.pos, a.ldp, a.faces[f * 3], a.faces[f * 3 + 1], a.faces[f * 3 + 2], n, e1, e2, len);
                double gnrm[3], dot = 0.0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double df = (double)n[d] - tget(a.tfn, a.t64, f * 3 + d);
                    gnrm[d] = df > 0.0 ? kf : (df < 0.0 ? -kf : 0.0);
                    dot += (double)n[d] * gnrm[d];
                }
                double dc[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) dc[d] = (gnrm[d] - (double)n[d] * dot) / (double)len;
                const double de1[3] = {e2[1] * dc[2] - e2[2] * dc[1], e2[2] * dc[0] - e2[0] * dc[2], e2[0] * dc[1] - e2[1] * dc[0]};
                const double de2[3] = {dc[1] * e1[2] - dc[2] * e1[1], dc[2] * e1[0] - dc[0] * e1[2], dc[0] * e1[1] - dc[1] * e1[0]};
#pragma unroll
                for (int d = 0; d < 3; ++d) g[d] += corner == 0 ? -(de1[d] + de2[d]) : (corner == 1 ? de1[d] : de2[d]);
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) b.dpos[v * b.lddpos + d] = (float)g[d];
    }
}

// ---------------- vertex -> face-corner incidence ----------------
__global__ void k_inc_count(const int64_t* __restrict__ faces, int64_t nf, int64_t n, int32_t* __restrict__ cnt, int32_t* __restrict__ err) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nf * 3; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = faces[i];
        if (v < 0 || v >= n) { *err = 1; continue; }
        atomicAdd(&cnt[v], 1);
    }
}
__global__ void k_inc_fill(const int64_t* __restrict__ faces, int64_t nf, int64_t n, const int32_t* __restrict__ rowptr,
                           int32_t* __restrict__ cursor, int32_t* __restrict__ inc) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nf * 3; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = faces[i];
        if (v < 0 || v >= n) continue;
        inc[rowptr[v] + atomicAdd(&cursor[v], 1)] = (int32_t)i;
    }
}
// ascending order inside each row (rows are short: insertion sort by one thread) -> deterministic backward sums
__global__ void k_inc_sort(const int32_t* __restrict__ rowptr, int64_t n, int32_t* __restrict__ inc) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        const int s = rowptr[v], t = rowptr[v + 1];
        for (int i = s + 1; i < t; ++i) {
            const int32_t key = inc[i];
            int j = i - 1;
            while (j >= s && inc[j] > key) { inc[j + 1] = inc[j]; --j; }
            inc[j + 1] = key;
        }
    }
}

// ---------------- Laplacian loss:  d_v = pos_v - (sum_{j in N(v)} pos_j) / deg_v ----------------
__global__ void __launch_bounds__(kLossThreads) k_lap_fwd(const float* __restrict__ pos, int64_t ldp, int64_t n, const int32_t* __restrict__ rowptr,
                                                          const int2* __restrict__ edges, float* __restrict__ diff, double* __restrict__ partials) {
    __shared__ double sh[kLossThreads];
    double sq = 0.0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        float s[3] = {0.f, 0.f, 0.f};
        const int b = rowptr[v], e = rowptr[v + 1];
        for (int p = b; p < e; ++p) {
            const int64_t j = edges[p].x;
#pragma unroll
            for (int d = 0; d < 3; ++d) s[d] += __ldg(pos + j * ldp + d);
        }
        const float deg = (float)(e - b);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float dv = __ldg(pos + v * ldp + d) - s[d] / deg;
            diff[v * 3 + d] = dv;
            sq += (double)dv * (double)dv;
        }
    }
    sq = block_sum(sq, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = sq;
}
__global__ void k_lap_finalize(const double* __restrict__ partials, int rows, int64_t n, double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += partials[r];
    out[0] = (double)sqrtf((float)(s / (double)n) + 1.0e-12f);
}
// grad_v = k * (d_v - sum_{j: v in N(j)} d_j / deg_j),  k = g / (n * loss);  rowptr_t/edges_t = by-source CSR
__global__ void __launch_bounds__(kLossThreads) k_lap_bwd(const float* __restrict__ diff, int64_t n, const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ rowptr_t, const int2* __restrict__ edges_t,
                                                          const double* __restrict__ out, const double* __restrict__ grad,
                                                          float* __restrict__ dpos, int64_t lddpos) {
    const double k = grad[0] / ((double)n * out[0]);
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        double g[3] = {diff[v * 3], diff[v * 3 + 1], diff[v * 3 + 2]};
        for (int p = rowptr_t[v]; p < rowptr_t[v + 1]; ++p) {
            const int64_t j = edges_t[p].x;
            const double dj = (double)(rowptr[j + 1] - rowptr[j]);
#pragma unroll
            for (int d = 0; d < 3; ++d) g[d] -= (double)diff[j * 3 + d] / dj;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) dpos[v * lddpos + d] = (float)(k * g[d]);
    }
}

static int loss_grid(int64_t work) {
    int64_t g = ceil_div(work > 0 ? work : 1, kLossThreads);
    int64_t cap = (int64_t)num_sms() * 4;
    if (cap > kLossRowsMax) cap = kLossRowsMax;
    return (int)(g < cap ? g : cap);
}

}  // namespace sgb

using namespace sgb;

extern "C" int sgb_loss_partial_rows(void) { return kLossRowsMax; }

extern "C" size_t sgb_incidence_build_workspace_bytes(int64_t nf, int64_t n) {
    if (nf < 0 || n < 0) return 0;
    return align_up((size_t)(n + 1) * 4, 256) * 2 + align_up((size_t)(ceil_div(n > 0 ? n : 1, 4096) + 1) * 4, 256);
}

extern "C" int sgb_incidence_build(const int64_t* faces, int64_t nf, int64_t n, int32_t* rowptr, int32_t* inc, int32_t* err_flag,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(nf >= 0 && n >= 0 && nf * 3 < (int64_t)0x7fffffff && n < (int64_t)0x7fffffff, "sgb_incidence_build: bad sizes");
    SGB_CHECK_ARG(rowptr && err_flag && (nf == 0 || (faces && inc)), "sgb_incidence_build: null pointer");
    const size_t need = sgb_incidence_build_workspace_bytes(nf, n);
    if (!workspace || workspace_bytes < need) {
        set_error("sgb_incidence_build: workspace %zu < required %zu", workspace_bytes, need);
        return SGB_ENOSPC;
    }
    SGB_CUDA(cudaMemsetAsync(workspace, 0, need, stream));
    SGB_CUDA(cudaMemsetAsync(err_flag, 0, sizeof(int32_t), stream));
    if (n == 0) {
        SGB_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int32_t), stream));
        return SGB_OK;
    }
    char* w = reinterpret_cast<char*>(workspace);
    int32_t* cnt = reinterpret_cast<int32_t*>(w);
    int32_t* cursor = reinterpret_cast<int32_t*>(w + align_up((size_t)(n + 1) * 4, 256));
    int32_t* tsum = reinterpret_cast<int32_t*>(w + 2 * align_up((size_t)(n + 1) * 4, 256));
    const int threads = 256;
    const int fgrid = (int)min64(ceil_div(nf * 3 > 0 ? nf * 3 : 1, threads), (int64_t)num_sms() * 16);
    const int ngrid = (int)min64(ceil_div(n, threads), (int64_t)num_sms() * 16);
    if (nf > 0) {
        k_inc_count<<<fgrid, threads, 0, stream>>>(faces, nf, n, cnt, err_flag);
        SGB_CHECK_LAUNCH("k_inc_count");
    }
    int rc = exclusive_scan_i32(cnt, n, tsum, rowptr, stream);
    if (rc != SGB_OK) return rc;
    if (nf > 0) {
        k_inc_fill<<<fgrid, threads, 0, stream>>>(faces, nf, n, rowptr, cursor, inc);
        SGB_CHECK_LAUNCH("k_inc_fill");
        k_inc_sort<<<ngrid, threads, 0, stream>>>(rowptr, n, inc);
        SGB_CHECK_LAUNCH("k_inc_sort");
    }
    return SGB_OK;
}

extern "C" int sgb_step_loss_fwd(const float* pos, int64_t ldp, int64_t n, const void* target_pos, int target_f64, const uint8_t* vmask,
                                 const int64_t* faces, int64_t nf, const void* target_fn, const uint8_t* fmask, float* fn_out,
                                 double* partials, double* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(pos && ldp >= 3 && n >= 0 && nf >= 0 && partials && out, "sgb_step_loss_fwd: bad argument");
    SGB_CHECK_ARG(nf == 0 || faces, "sgb_step_loss_fwd: faces missing");
    LossArgs a{pos, ldp, n, target_pos, target_f64, vmask, nf > 0 ? faces : nullptr, nf, target_fn, fmask};
    const int grid = loss_grid(n > nf ? n : nf);
    k_step_loss_fwd<<<grid, kLossThreads, 0, stream>>>(a, fn_out, partials);
    SGB_CHECK_LAUNCH("k_step_loss_fwd");
    k_step_loss_finalize<<<1, kLossThreads, 0, stream>>>(partials, grid, target_f64, out);
    SGB_CHECK_LAUNCH("k_step_loss_finalize");
    return SGB_OK;
}

extern "C" int sgb_step_loss_bwd(const float* pos, int64_t ldp, int64_t n, const void* target_pos, int target_f64, const uint8_t* vmask,
                                 const int64_t* faces, int64_t nf, const void* target_fn, const uint8_t* fmask,
                                 const int32_t* inc_rowptr, const int32_t* inc, const double* out, const double* grads,
                                 float* dpos, int64_t lddpos, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(pos && ldp >= 3 && n >= 0 && out && grads && dpos && lddpos >= 3, "sgb_step_loss_bwd: bad argument");
    SGB_CHECK_ARG(nf == 0 || !target_fn || (faces && inc_rowptr && inc), "sgb_step_loss_bwd: faces / incidence missing");
    if (n == 0) return SGB_OK;
    LossBwdArgs b{{pos, ldp, n, target_pos, target_f64, vmask, nf > 0 ? faces : nullptr, nf, target_fn, fmask}, inc_rowptr, inc, out, grads, dpos, lddpos};
    k_step_loss_bwd<<<loss_grid(n), kLossThreads, 0, stream>>>(b);
    SGB_CHECK_LAUNCH("k_step_loss_bwd");
    return SGB_OK;
}

extern "C" int sgb_lap_loss_fwd(const float* pos, int64_t ldp, int64_t n, const int32_t* rowptr, const sgb_edge_t* edges, float* diff,
                                double* partials, double* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(pos && ldp >= 3 && n > 0 && rowptr && edges && diff && partials && out, "sgb_lap_loss_fwd: bad argument");
    const int grid = loss_grid(n);
    k_lap_fwd<<<grid, kLossThreads, 0, stream>>>(pos, ldp, n, rowptr, reinterpret_cast<const int2*>(edges), diff, partials);
    SGB_CHECK_LAUNCH("k_lap_fwd");
    k_lap_finalize<<<1, 32, 0, stream>>>(partials, grid, n, out);
    SGB_CHECK_LAUNCH("k_lap_finalize");
    return SGB_OK;
}

extern "C" int sgb_lap_loss_bwd(const float* diff, int64_t n, const int32_t* rowptr, const int32_t* rowptr_t, const sgb_edge_t* edges_t,
                                const double* out, const double* grad, float* dpos, int64_t lddpos, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(diff && n > 0 && rowptr && rowptr_t && edges_t && out && grad && dpos && lddpos >= 3, "sgb_lap_loss_bwd: bad argument");
    k_lap_bwd<<<loss_grid(n), kLossThreads, 0, stream>>>(diff, n, rowptr, rowptr_t, reinterpret_cast<const int2*>(edges_t), out, grad, dpos, lddpos);
    SGB_CHECK_LAUNCH("k_lap_bwd");
    return SGB_OK;
}
