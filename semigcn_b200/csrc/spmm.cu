// Row-parallel SpMM over the mesh vertex graph:  Y = alpha * (S f(X)) + beta * ADDEND + bias.
//
// Replaces PyG's MessagePassing.propagate (x.index_select(0,row) -> norm.view(-1,1)*x_j ->
// scatter_add at col), SURVEY.md §3.2 / A.1 step 3 / A.2 step 3, forward and -- on the
// transpose CSR -- backward.  No [nnz, C] message tensor, no atomics, edge weights recomputed
// from dis (never stored), neighbour indices staged in registers and broadcast by shuffles.
//
// Mapping: one sub-warp of LPV lanes per vertex, each lane owns VEC contiguous channels per
// iteration (128-bit loads when C % 4 == 0), ITERS iterations cover up to LPV*VEC*ITERS
// channels per pass; persistent grid (multiple of the SM count) with a grid-stride loop over
// vertices so consecutive vertices are in flight together (their neighbourhoods overlap ->
// L1/L2 hits; DRAM sees X once).  Up to 8 independent 128-bit gathers in flight per lane.
//
// Arithmetic order is the reference's: per row, messages in CSR (= edge) order, each
// msg = fl(w * x_j) with w = fl(dis_j * dis_i), acc = fl(acc + msg); then the self-loop
// term(s): GCN  acc += fl(fl(dis_i*dis_i) * x_i);  CHEB  acc = fl(fl(acc + x_i) - x_i).
// => bit-identical to the CPU oracle on the same inputs.
//
// HBM roofline: algorithmic bytes = 2*N*C*4 + (nnz + 2N + 1)*4.
#include "common.cuh"

namespace sgb {

constexpr int kSpmmThreads = 256;

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        float4 t = ldg4(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float* p) const { st4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <>
struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};

struct SpmmArgs {
    const int32_t* rowptr;
    const int32_t* colidx;
    const float* dis;
    int mode;
    const float* x;
    int64_t ldx;
    int64_t n;
    int c;
    const float* in_mean;
    const float* in_scale;
    const float* in_shift;
    float slope;
    float alpha;
    const float* addend;
    int64_t ld_addend;
    float beta;
    const float* bias;
    float* y;
    int64_t ldy;
    float* stat_partials;
};

constexpr int kSpmmVpc = 64;      // vertices per chunk (contiguous ids); a CTA owns a contiguous range of chunks
constexpr int kSpmmEcap = 1024;   // neighbour indices staged in shared memory per chunk (mesh: ~6 * 64 = 384)

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Pipeline per CTA: while chunk c is being gathered, the CSR slice of chunk c+1 (colidx) and the row
// pointers of chunk c+2 are already in flight (cp.async into the other half of the double buffer).
// Per vertex there is exactly ONE round of dependent global loads: the neighbour rows, the self row
// and the dis[] entries are all issued together (indices come from shared memory).
template <int LPV, int VEC, int ITERS, bool PRO, bool STATS>
__global__ void __launch_bounds__(kSpmmThreads, (VEC == 4) ? 2 : 3) k_spmm(const SpmmArgs a) {
    constexpr int NB = 6;                                   // neighbour rows loaded per round (+ the self row in round 0)
    constexpr int VPI = (ITERS == 1) ? 2 : 1;               // vertices in flight per sub-warp
    constexpr int CH = LPV * VEC * ITERS;                   // channels per pass
    constexpr int GROUPS = kSpmmThreads / LPV;              // sub-warps per CTA
    const int l = threadIdx.x & (LPV - 1);
    const int grp = threadIdx.x / LPV;
    constexpr bool pro = PRO;        // compile-time: the BatchNorm vectors / statistics accumulators cost ~30 registers
    constexpr bool stats = STATS;
    const bool self_slot = a.mode != SGB_MODE_ADJ;
    __shared__ float red[2][kSpmmThreads * VEC * ITERS];
    __shared__ float redn[kSpmmThreads];
    __shared__ int s_rowptr[2][kSpmmVpc + 1];
    __shared__ int s_col[2][kSpmmEcap];
    const int64_t nchunks = (a.n + kSpmmVpc - 1) / kSpmmVpc;
    const int64_t cpc = (nchunks + gridDim.x - 1) / gridDim.x;
    const int64_t cbeg = (int64_t)blockIdx.x * cpc;
    const int64_t cend = min64(nchunks, cbeg + cpc);

    auto stage_rowptr = [&](int64_t chunk, int buf) {      // async; chunk may be past the end
        if (chunk < cend) {
            const int64_t v0 = chunk * kSpmmVpc;
            const int nv = (int)min64(kSpmmVpc, a.n - v0);
            for (int i = threadIdx.x; i <= nv; i += kSpmmThreads) cp_async4(&s_rowptr[buf][i], a.rowptr + v0 + i);
        }
    };
    auto stage_col = [&](int64_t chunk, int buf) {         // needs s_rowptr[buf] of that chunk to be visible
        if (chunk < cend) {
            const int64_t v0 = chunk * kSpmmVpc;
            const int nv = (int)min64(kSpmmVpc, a.n - v0);
            const int e0 = s_rowptr[buf][0], ne = s_rowptr[buf][nv] - e0;
            if (ne <= kSpmmEcap)
                for (int i = threadIdx.x; i < ne; i += kSpmmThreads) cp_async4(&s_col[buf][i], a.colidx + e0 + i);
        }
    };

    for (int c0 = 0; c0 < a.c; c0 += CH) {
        int ch[ITERS];
        bool act[ITERS];
        Vec<VEC> mu[ITERS], sc[ITERS], sh[ITERS], bs[ITERS];
#pragma unroll
        for (int t = 0; t < ITERS; ++t) {
            ch[t] = c0 + (t * LPV + l) * VEC;
            act[t] = ch[t] < a.c;
#pragma unroll
            for (int q = 0; q < VEC; ++q) { mu[t].v[q] = 0.f; sc[t].v[q] = 1.f; sh[t].v[q] = 0.f; bs[t].v[q] = 0.f; }
            if (act[t]) {
                if (pro) { mu[t].load(a.in_mean + ch[t]); sc[t].load(a.in_scale + ch[t]); sh[t].load(a.in_shift + ch[t]); }
                if (a.bias) bs[t].load(a.bias + ch[t]);
            }
        }
        // statistics: pivot-shifted sums per thread (pivot = first value seen), see common.cuh
        float s1[ITERS][VEC], s2[ITERS][VEC], pv[ITERS][VEC];
        float nseen = 0.f;
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
#pragma unroll
            for (int q = 0; q < VEC; ++q) { s1[t][q] = 0.f; s2[t][q] = 0.f; pv[t][q] = 0.f; }

        // ---- pipeline prologue: rowptr(cbeg) -> colidx(cbeg) + rowptr(cbeg+1)
        __syncthreads();
        stage_rowptr(cbeg, 0);
        cp_async_commit_wait_all();
        __syncthreads();
        stage_col(cbeg, 0);
        stage_rowptr(cbeg + 1, 1);

        for (int64_t chunk = cbeg; chunk < cend; ++chunk) {
            const int buf = (int)((chunk - cbeg) & 1);
            cp_async_commit_wait_all();
            __syncthreads();                                  // chunk's colidx + next chunk's rowptr have landed; previous gather done
            stage_col(chunk + 1, buf ^ 1);
            // (rowptr of chunk+2 goes into the buffer this chunk is reading: issued after the gather below)
            const int64_t v0 = chunk * kSpmmVpc;
            const int nv = (int)min64(kSpmmVpc, a.n - v0);
            const int e0 = s_rowptr[buf][0];
            const bool staged = (s_rowptr[buf][nv] - e0) <= kSpmmEcap;
            const int* scol = s_col[buf];
            const int* srp = s_rowptr[buf];

            for (int vb = grp; vb < nv; vb += GROUPS * VPI) {
                int vi[VPI], kbeg[VPI], kend[VPI];
                float di[VPI];
                Vec<VEC> acc[VPI][ITERS], xself[VPI][ITERS];
#pragma unroll
                for (int u = 0; u < VPI; ++u) {
                    vi[u] = vb + u * GROUPS;
                    const bool vok = vi[u] < nv;
                    kbeg[u] = vok ? srp[vi[u]] - e0 : 0;
                    kend[u] = vok ? srp[vi[u] + 1] - e0 : -1;   // -1 => no vertex
                    di[u] = vok ? __ldg(a.dis + v0 + vi[u]) : 0.f;
#pragma unroll
                    for (int t = 0; t < ITERS; ++t)
#pragma unroll
                        for (int q = 0; q < VEC; ++q) { acc[u][t].v[q] = 0.f; xself[u][t].v[q] = 0.f; }
                }
                for (int r = 0;; r += NB) {
                    bool more = false;
                    Vec<VEC> xv[VPI][NB][ITERS];
                    float dj[VPI][NB];
#pragma unroll
                    for (int u = 0; u < VPI; ++u) {
                        if (r == 0 && self_slot && kend[u] >= 0) {          // self row rides along with the first round
#pragma unroll
                            for (int t = 0; t < ITERS; ++t)
                                if (act[t]) xself[u][t].load(a.x + (v0 + vi[u]) * a.ldx + ch[t]);
                        }
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const int k = kbeg[u] + r + b;
                            dj[u][b] = 0.f;
                            if (k < kend[u]) {
                                const int64_t j = staged ? scol[k] : __ldg(a.colidx + e0 + k);
                                dj[u][b] = __ldg(a.dis + j);
#pragma unroll
                                for (int t = 0; t < ITERS; ++t)
                                    if (act[t]) xv[u][b][t].load(a.x + j * a.ldx + ch[t]);
                            }
                        }
                        more |= (kbeg[u] + r + NB) < kend[u];
                    }
#pragma unroll
                    for (int u = 0; u < VPI; ++u) {
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const int k = kbeg[u] + r + b;
                            if (k < kend[u]) {
                                // w = fl(dis[row] * dis[col]) (A.1 step 1); CHEB negates (exact); ADJ = 1
                                float w = __fmul_rn(dj[u][b], di[u]);
                                if (a.mode == SGB_MODE_CHEB) w = -w;
                                if (a.mode == SGB_MODE_ADJ) w = 1.f;
#pragma unroll
                                for (int t = 0; t < ITERS; ++t)
                                    if (act[t]) {
#pragma unroll
                                        for (int q = 0; q < VEC; ++q) {
                                            float xx = xv[u][b][t].v[q];
                                            if (pro) xx = bn_lrelu(xx, mu[t].v[q], sc[t].v[q], sh[t].v[q], a.slope);
                                            acc[u][t].v[q] = __fadd_rn(acc[u][t].v[q], __fmul_rn(w, xx));
                                        }
                                    }
                            }
                        }
                    }
                    if (!more) break;
                }
#pragma unroll
                for (int u = 0; u < VPI; ++u) {
                    if (vi[u] >= nv) continue;
                    const int64_t v = v0 + vi[u];
#pragma unroll
                    for (int t = 0; t < ITERS; ++t) {
                        if (!act[t]) continue;
                        Vec<VEC> out = acc[u][t];
                        if (self_slot) {
                            const float wii = __fmul_rn(di[u], di[u]);
#pragma unroll
                            for (int q = 0; q < VEC; ++q) {
                                float xx = xself[u][t].v[q];
                                if (pro) xx = bn_lrelu(xx, mu[t].v[q], sc[t].v[q], sh[t].v[q], a.slope);
                                if (a.mode == SGB_MODE_GCN) {
                                    out.v[q] = __fadd_rn(out.v[q], __fmul_rn(wii, xx));
                                } else {   // CHEB: the (+1, -1) loop pair of ChebConv.__norm__, not cancelled
                                    out.v[q] = __fadd_rn(__fadd_rn(out.v[q], xx), -xx);
                                }
                            }
                        }
                        if (a.addend) {
                            Vec<VEC> ad;
                            ad.load(a.addend + v * a.ld_addend + ch[t]);
#pragma unroll
                            for (int q = 0; q < VEC; ++q)
                                out.v[q] = __fadd_rn(__fmul_rn(a.alpha, out.v[q]), __fmul_rn(a.beta, ad.v[q]));
                        } else if (a.alpha != 1.f) {
#pragma unroll
                            for (int q = 0; q < VEC; ++q) out.v[q] = __fmul_rn(a.alpha, out.v[q]);
                        }
                        if (a.bias) {
#pragma unroll
                            for (int q = 0; q < VEC; ++q) out.v[q] = __fadd_rn(out.v[q], bs[t].v[q]);
                        }
                        out.store(a.y + v * a.ldy + ch[t]);
                        if (stats) {
#pragma unroll
                            for (int q = 0; q < VEC; ++q) {
                                if (nseen == 0.f) pv[t][q] = out.v[q];
                                const float dv = out.v[q] - pv[t][q];
                                s1[t][q] += dv;
                                s2[t][q] = fmaf(dv, dv, s2[t][q]);
                            }
                        }
                    }
                    nseen += 1.f;
                }
            }
            __syncthreads();                                  // everyone is done reading s_rowptr[buf]
            stage_rowptr(chunk + 2, buf);
        }
        cp_async_commit_wait_all();
        if (stats) {   // fixed-order block merge (Chan) -> partials[blockIdx.x][3][c] = (count, mean, M2)
#pragma unroll
            for (int t = 0; t < ITERS; ++t)
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    const int cl = (t * LPV + l) * VEC + q;     // channel within the pass
                    const Moments mo = from_shifted(nseen, pv[t][q], s1[t][q], s2[t][q]);
                    red[0][grp * CH + cl] = mo.mean;
                    red[1][grp * CH + cl] = mo.m2;
                }
            if (l == 0) redn[grp] = nseen;
            __syncthreads();
            for (int cl = threadIdx.x; cl < CH; cl += kSpmmThreads) {
                if (c0 + cl < a.c) {
                    Moments acc{0.f, 0.f, 0.f};
                    for (int g = 0; g < GROUPS; ++g) acc = merge(acc, Moments{redn[g], red[0][g * CH + cl], red[1][g * CH + cl]});
                    a.stat_partials[((int64_t)blockIdx.x * 3 + 0) * a.c + c0 + cl] = acc.n;
                    a.stat_partials[((int64_t)blockIdx.x * 3 + 1) * a.c + c0 + cl] = acc.mean;
                    a.stat_partials[((int64_t)blockIdx.x * 3 + 2) * a.c + c0 + cl] = acc.m2;
                }
            }
            __syncthreads();
        }
    }
}

struct SpmmCfg {
    int lpv, vec, iters;
};

static SpmmCfg pick_cfg(int c, bool aligned) {
    SpmmCfg k;
    if (c % 4 == 0 && aligned) {
        k.vec = 4;
        int lanes = c / 4;
        int p = 1;
        while (p < lanes && p < 32) p <<= 1;
        k.lpv = p;
        k.iters = lanes > 32 ? 2 : 1;
    } else {
        k.vec = 1;
        k.lpv = c <= 4 ? 4 : 32;
        k.iters = 1;
    }
    return k;
}

static int spmm_grid(int64_t n, const SpmmCfg& k) {
    int64_t need = ceil_div(n > 0 ? n : 1, kSpmmVpc);
    int per_sm = (k.vec == 4) ? 2 : 3;
    int64_t cap = (int64_t)num_sms() * per_sm;
    return (int)(need < cap ? need : cap);
}

}  // namespace sgb

extern "C" int sgb_spmm_stat_rows(int64_t n, int c) {
    if (n < 0 || c <= 0) return 0;
    // alignment of x/y is not known here: the row count must not depend on it, so both
    // candidate configurations are sized and the larger grid is reported.
    int g1 = sgb::spmm_grid(n, sgb::pick_cfg(c, true));
    int g2 = sgb::spmm_grid(n, sgb::pick_cfg(c, false));
    return g1 > g2 ? g1 : g2;
}

extern "C" int sgb_spmm(const int32_t* rowptr, const int32_t* colidx, const float* dis, int mode,
                        const float* x, int64_t ldx, int64_t n, int c,
                        const float* in_mean, const float* in_scale, const float* in_shift, float slope,
                        float alpha, const float* addend, int64_t ld_addend, float beta,
                        const float* bias, float* y, int64_t ldy, float* stat_partials, void* stream_) {
    using namespace sgb;
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(n >= 0 && c > 0, "sgb_spmm: bad shape n=%lld c=%d", (long long)n, c);
    SGB_CHECK_ARG(mode == SGB_MODE_GCN || mode == SGB_MODE_CHEB || mode == SGB_MODE_ADJ, "sgb_spmm: bad mode %d", mode);
    SGB_CHECK_ARG(rowptr && dis && x && y, "sgb_spmm: null pointer");
    SGB_CHECK_ARG(ldx >= c && ldy >= c && (!addend || ld_addend >= c), "sgb_spmm: leading dimension < c");
    SGB_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr) && (in_scale == nullptr) == (in_mean == nullptr),
                  "sgb_spmm: in_mean / in_scale / in_shift must come together");
    SGB_CHECK_ARG(x != y, "sgb_spmm: in-place aggregation is not supported");
    if (n == 0) return SGB_OK;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    bool aligned = al16(x) && al16(y) && ldx % 4 == 0 && ldy % 4 == 0 && (!addend || (al16(addend) && ld_addend % 4 == 0)) &&
                   (!bias || al16(bias)) && (!in_scale || (al16(in_scale) && al16(in_shift) && al16(in_mean)));
    SpmmCfg k = pick_cfg(c, aligned);
    int grid = spmm_grid(n, k);
    if (stat_partials) {
        // rows the caller sized for; unused rows must read as zero
        int rows = sgb_spmm_stat_rows(n, c);
        if (rows > grid)
            SGB_CUDA(cudaMemsetAsync(stat_partials + (size_t)grid * 3 * c, 0, (size_t)(rows - grid) * 3 * c * sizeof(float), stream));
    }
    SpmmArgs a{rowptr, colidx, dis, mode, x, ldx, n, c, in_mean, in_scale, in_shift, slope, alpha, addend, ld_addend, beta, bias, y, ldy, stat_partials};
    const bool pro = in_scale != nullptr, st = stat_partials != nullptr;
#define SGB_SPMM_CASE(L, V, I)                                                                  \
    if (k.lpv == L && k.vec == V && k.iters == I) {                                             \
        if (pro && st) k_spmm<L, V, I, true, true><<<grid, kSpmmThreads, 0, stream>>>(a);       \
        else if (pro) k_spmm<L, V, I, true, false><<<grid, kSpmmThreads, 0, stream>>>(a);       \
        else if (st) k_spmm<L, V, I, false, true><<<grid, kSpmmThreads, 0, stream>>>(a);        \
        else k_spmm<L, V, I, false, false><<<grid, kSpmmThreads, 0, stream>>>(a);               \
        SGB_CHECK_LAUNCH("k_spmm");                                                             \
        return SGB_OK;                                                                          \
    }
    SGB_SPMM_CASE(1, 4, 1)
    SGB_SPMM_CASE(2, 4, 1)
    SGB_SPMM_CASE(4, 4, 1)
    SGB_SPMM_CASE(8, 4, 1)
    SGB_SPMM_CASE(16, 4, 1)
    SGB_SPMM_CASE(32, 4, 1)
    SGB_SPMM_CASE(32, 4, 2)
    SGB_SPMM_CASE(4, 1, 1)
    SGB_SPMM_CASE(32, 1, 1)
#undef SGB_SPMM_CASE
    set_error("sgb_spmm: no kernel configuration for c=%d", c);
    return SGB_ENOTSUP;
}
