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
#include <stdlib.h>

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
    const int2* edges;                // [nnz] (col, float bits of the edge weight), CSR order
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
    const float* xg;                  // ghost rows (vertex-partitioned mode): neighbour ids >= split read xg[(id - split)]
    int64_t ldxg;
    int32_t split;                    // INT32_MAX when there are no ghost rows
    int strided;                      // 1: CTA b takes chunks b, b+grid, ... (co-resident CTAs sweep one window of X -> L2 reuse)
};

// vertices per chunk / staged edges per chunk as a function of the sub-warp width: every sub-warp of the CTA
// gets at least one vertex per chunk
__host__ __device__ constexpr int spmm_vpc(int lpv) { return (kSpmmThreads / lpv) > 64 ? (kSpmmThreads / lpv) : 64; }
__host__ __device__ constexpr int spmm_ecap(int lpv) { return spmm_vpc(lpv) * 12 < 1536 ? spmm_vpc(lpv) * 12 : 1536; }

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Pipeline per CTA: while chunk i is being gathered, the (col, weight) slice of chunk i+1 and the row
// pointers of chunk i+2 are already in flight (cp.async into the other half of the double buffer).
// Per vertex there is exactly ONE round of dependent global loads for degree <= NB: the neighbour rows
// and the self row are all issued together (indices and weights come from shared memory).
template <int LPV, int VEC, int ITERS, bool PRO, bool STATS>
#ifndef SGB_SPMM_MINB
#define SGB_SPMM_MINB 2
#endif
__global__ void __launch_bounds__(kSpmmThreads, (VEC == 4) ? SGB_SPMM_MINB : 3) k_spmm(const SpmmArgs a) {
#ifndef SGB_SPMM_NB
#define SGB_SPMM_NB 6
#endif
    constexpr int NB = (VEC * ITERS >= 8) ? SGB_SPMM_NB : 8;  // neighbour rows in flight per round (register budget)
    constexpr int CH = LPV * VEC * ITERS;                   // channels per pass
    constexpr int GROUPS = kSpmmThreads / LPV;              // sub-warps per CTA
    constexpr int VPC = spmm_vpc(LPV);
    constexpr int ECAP = spmm_ecap(LPV);
    const int l = threadIdx.x & (LPV - 1);
    const int grp = threadIdx.x / LPV;
    __shared__ float red[STATS ? 2 : 1][STATS ? kSpmmThreads * VEC * ITERS : 1];
    __shared__ float redn[STATS ? kSpmmThreads : 1];
    __shared__ int s_rowptr[2][VPC + 1];
    __shared__ int2 s_edge[2][ECAP];
    const int64_t nchunks = (a.n + VPC - 1) / VPC;
    // chunk schedule: chunk(i) = cfirst + i * cstep for i < ccount
    const int64_t cpc = (nchunks + gridDim.x - 1) / gridDim.x;
    const int64_t cstep = a.strided ? (int64_t)gridDim.x : 1;
    const int64_t cfirst = a.strided ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * cpc;
    const int64_t ccount = a.strided ? (nchunks > blockIdx.x ? (nchunks - 1 - blockIdx.x) / gridDim.x + 1 : 0)
                                     : (min64(nchunks, cfirst + cpc) > cfirst ? min64(nchunks, cfirst + cpc) - cfirst : 0);

    auto stage_rowptr = [&](int64_t ci, int buf) {         // async; ci (schedule position) may be past the end
        if (ci < ccount) {
            const int64_t v0 = (cfirst + ci * cstep) * VPC;
            const int nv = (int)min64(VPC, a.n - v0);
            for (int i = threadIdx.x; i <= nv; i += kSpmmThreads) cp_async4(&s_rowptr[buf][i], a.rowptr + v0 + i);
        }
    };
    auto stage_edges = [&](int64_t ci, int buf) {          // needs s_rowptr[buf] of that chunk to be visible
        if (ci < ccount) {
            const int64_t v0 = (cfirst + ci * cstep) * VPC;
            const int nv = (int)min64(VPC, a.n - v0);
            const int e0 = s_rowptr[buf][0], ne = s_rowptr[buf][nv] - e0;
            if (ne <= ECAP)
                for (int i = threadIdx.x; i < ne; i += kSpmmThreads) cp_async8(&s_edge[buf][i], a.edges + e0 + i);
        }
    };

    for (int c0 = 0; c0 < a.c; c0 += CH) {
        int ch[ITERS];
        bool act[ITERS];
        Vec<VEC> mu[ITERS], sc[ITERS], sh[ITERS];
#pragma unroll
        for (int t = 0; t < ITERS; ++t) {
            ch[t] = c0 + (t * LPV + l) * VEC;
            act[t] = ch[t] < a.c;
#pragma unroll
            for (int q = 0; q < VEC; ++q) { mu[t].v[q] = 0.f; sc[t].v[q] = 1.f; sh[t].v[q] = 0.f; }
            if (act[t]) {
                if (PRO) { mu[t].load(a.in_mean + ch[t]); sc[t].load(a.in_scale + ch[t]); sh[t].load(a.in_shift + ch[t]); }
            }
        }
        const float* xl = a.x + ch[0];                       // this lane's first channel; iteration t adds t * LPV * VEC
        // statistics: pivot-shifted sums per thread (pivot = first value seen), see common.cuh
        float s1[ITERS][VEC], s2[ITERS][VEC], pv[ITERS][VEC];
        float nseen = 0.f;
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
#pragma unroll
            for (int q = 0; q < VEC; ++q) { s1[t][q] = 0.f; s2[t][q] = 0.f; pv[t][q] = 0.f; }

        // ---- pipeline prologue: rowptr(0) -> edges(0) + rowptr(1)
        __syncthreads();
        stage_rowptr(0, 0);
        cp_async_commit_wait_all();
        __syncthreads();
        stage_edges(0, 0);
        stage_rowptr(1, 1);

        for (int64_t ci = 0; ci < ccount; ++ci) {
            const int buf = (int)(ci & 1);
            cp_async_commit_wait_all();
            __syncthreads();                                  // chunk's edges + next chunk's rowptr have landed; previous gather done
            stage_edges(ci + 1, buf ^ 1);
            // (rowptr of chunk ci+2 goes into the buffer this chunk is reading: issued after the gather below)
            const int64_t v0 = (cfirst + ci * cstep) * VPC;
            const int nv = (int)min64(VPC, a.n - v0);
            const int e0 = s_rowptr[buf][0];
            const bool staged = (s_rowptr[buf][nv] - e0) <= ECAP;
            const int2* sedge = s_edge[buf];
            const int* srp = s_rowptr[buf];

            // one instantiation per edge source (shared-memory slice, or global for chunks whose rows are too long to
            // stage): a per-neighbour select between the two would turn into a branch and serialise the gathers
            auto gather = [&](auto fetch_edge) {
                for (int vi = grp; vi < nv; vi += GROUPS) {
                    const int kb = srp[vi] - e0, ke = srp[vi + 1] - e0;
                    const int64_t v = v0 + vi;
                    Vec<VEC> acc[ITERS], xself[ITERS];
                    float di = 0.f;
    #pragma unroll
                    for (int t = 0; t < ITERS; ++t)
    #pragma unroll
                        for (int q = 0; q < VEC; ++q) { acc[t].v[q] = 0.f; xself[t].v[q] = 0.f; }
                    if (a.mode != SGB_MODE_ADJ) {                 // the self row rides along with the first round of gathers
                        if (a.mode == SGB_MODE_GCN) di = __ldg(a.dis + v);
    #pragma unroll
                        for (int t = 0; t < ITERS; ++t)
                            if (act[t]) xself[t].load(xl + v * a.ldx + t * LPV * VEC);
                    }
                    for (int r = kb; r < ke; r += NB) {
                        Vec<VEC> xv[NB][ITERS];
                        float w[NB];
    #pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            if (r + b < ke) {
                                const int2 ed = fetch_edge(r + b);
                                w[b] = __int_as_float(ed.y);
                                // owned rows live in x, halo (ghost) rows of the partitioned mode in xg
                            const float* xr = ed.x < a.split ? xl + (int64_t)ed.x * a.ldx : a.xg + ch[0] + (int64_t)(ed.x - a.split) * a.ldxg;
    #pragma unroll
                                for (int t = 0; t < ITERS; ++t)
                                    if (act[t]) xv[b][t].load(xr + t * LPV * VEC);
                            }
                        }
    #pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            if (r + b < ke) {
                                // msg = fl(w * x_j), acc = fl(acc + msg): PyG's message / aggregate op order (A.1 step 3)
    #pragma unroll
                                for (int t = 0; t < ITERS; ++t)
                                    if (act[t]) {
    #pragma unroll
                                        for (int q = 0; q < VEC; ++q) {
                                            float xx = xv[b][t].v[q];
                                            if (PRO) xx = bn_lrelu(xx, mu[t].v[q], sc[t].v[q], sh[t].v[q], a.slope);
                                            acc[t].v[q] = __fadd_rn(acc[t].v[q], __fmul_rn(w[b], xx));
                                        }
                                    }
                            }
                        }
                    }
                    const float wii = __fmul_rn(di, di);
    #pragma unroll
                    for (int t = 0; t < ITERS; ++t) {
                        if (!act[t]) continue;
                        Vec<VEC> out = acc[t];
                        if (a.mode != SGB_MODE_ADJ) {
    #pragma unroll
                            for (int q = 0; q < VEC; ++q) {
                                float xx = xself[t].v[q];
                                if (PRO) xx = bn_lrelu(xx, mu[t].v[q], sc[t].v[q], sh[t].v[q], a.slope);
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
                        if (a.bias) {                             // L1-resident; not worth 4*ITERS registers across the gather loop
                            Vec<VEC> bs;
                            bs.load(a.bias + ch[t]);
    #pragma unroll
                            for (int q = 0; q < VEC; ++q) out.v[q] = __fadd_rn(out.v[q], bs.v[q]);
                        }
                        out.store(a.y + v * a.ldy + ch[t]);
                        if (STATS) {
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
            };
#ifndef SGB_SPMM_BRANCH
            gather([&](int k) { return staged ? sedge[k] : __ldg(a.edges + e0 + k); });
#else
            if (staged) gather([&](int k) { return sedge[k]; });
            else gather([&](int k) { return __ldg(a.edges + e0 + k); });
#endif
            __syncthreads();                                  // everyone is done reading s_rowptr[buf]
            stage_rowptr(ci + 2, buf);
        }
        cp_async_commit_wait_all();
        if (STATS) {   // fixed-order block merge (Chan) -> partials[blockIdx.x][3][c] = (count, mean, M2)
#pragma unroll
            for (int t = 0; t < ITERS; ++t)
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    const int cl = (t * LPV + l) * VEC + q;     // channel within the pass
                    const Moments mo = from_shifted(nseen, pv[t][q], s1[t][q], s2[t][q]);
                    red[0][grp * CH + cl] = mo.mean;
                    red[STATS ? 1 : 0][grp * CH + cl] = mo.m2;
                }
            if (l == 0) redn[grp] = nseen;
            __syncthreads();
            for (int cl = threadIdx.x; cl < CH; cl += kSpmmThreads) {
                if (c0 + cl < a.c) {
                    Moments acc{0.f, 0.f, 0.f};
                    for (int g = 0; g < GROUPS; ++g) acc = merge(acc, Moments{redn[g], red[0][g * CH + cl], red[STATS ? 1 : 0][g * CH + cl]});
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
        // 128-bit lanes, two per lane wherever the row is wide enough: halves the per-vertex index / address
        // instruction overhead (the kernel is issue-bound, not DRAM-bound, at one float4 per lane)
        k.vec = 4;
        const int lanes = c / 4;
        if (lanes == 1) { k.lpv = 1; k.iters = 1; }
        else {
            int p = 1;
            while (p * 2 < lanes && p < 32) p <<= 1;
            k.lpv = p;
            k.iters = 2;
        }
    } else {
        k.vec = 1;
        k.lpv = c <= 4 ? 4 : 32;
        k.iters = 1;
    }
    return k;
}

static int spmm_grid(int64_t n, const SpmmCfg& k) {
    int64_t need = ceil_div(n > 0 ? n : 1, spmm_vpc(k.lpv));
    int per_sm = (k.vec == 4) ? SGB_SPMM_MINB : 3;
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

extern "C" int sgb_spmm(const int32_t* rowptr, const sgb_edge_t* edges, const float* dis, int mode,
                        const float* x, int64_t ldx, int64_t n, int c,
                        const float* in_mean, const float* in_scale, const float* in_shift, float slope,
                        float alpha, const float* addend, int64_t ld_addend, float beta,
                        const float* bias, float* y, int64_t ldy, float* stat_partials, void* stream_) {
    return sgb_spmm_halo(rowptr, edges, dis, mode, x, ldx, n, c, nullptr, 0, n, in_mean, in_scale, in_shift, slope, alpha, addend, ld_addend,
                         beta, bias, y, ldy, stat_partials, stream_);
}

extern "C" int sgb_spmm_halo(const int32_t* rowptr, const sgb_edge_t* edges, const float* dis, int mode,
                             const float* x, int64_t ldx, int64_t n, int c, const float* x_ghost, int64_t ld_ghost, int64_t n_split,
                             const float* in_mean, const float* in_scale, const float* in_shift, float slope,
                             float alpha, const float* addend, int64_t ld_addend, float beta,
                             const float* bias, float* y, int64_t ldy, float* stat_partials, void* stream_) {
    using namespace sgb;
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(n >= 0 && c > 0, "sgb_spmm: bad shape n=%lld c=%d", (long long)n, c);
    SGB_CHECK_ARG(mode == SGB_MODE_GCN || mode == SGB_MODE_CHEB || mode == SGB_MODE_ADJ, "sgb_spmm: bad mode %d", mode);
    SGB_CHECK_ARG(rowptr && dis && x && y, "sgb_spmm: null pointer");
    SGB_CHECK_ARG((reinterpret_cast<uintptr_t>(edges) & 7) == 0, "sgb_spmm: edges must be 8-byte aligned");
    SGB_CHECK_ARG(ldx >= c && ldy >= c && (!addend || ld_addend >= c), "sgb_spmm: leading dimension < c");
    SGB_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr) && (in_scale == nullptr) == (in_mean == nullptr),
                  "sgb_spmm: in_mean / in_scale / in_shift must come together");
    SGB_CHECK_ARG(x != y, "sgb_spmm: in-place aggregation is not supported");
    SGB_CHECK_ARG(!x_ghost || (ld_ghost >= c && n_split >= 0 && n_split < (int64_t)0x7fffffff), "sgb_spmm_halo: bad ghost block");
    if (n == 0) return SGB_OK;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    bool aligned = al16(x) && al16(y) && ldx % 4 == 0 && ldy % 4 == 0 && (!x_ghost || (al16(x_ghost) && ld_ghost % 4 == 0)) && (!addend || (al16(addend) && ld_addend % 4 == 0)) &&
                   (!bias || al16(bias)) && (!in_scale || (al16(in_scale) && al16(in_shift) && al16(in_mean)));
    SpmmCfg k = pick_cfg(c, aligned);
    int grid = spmm_grid(n, k);
    if (stat_partials) {
        // rows the caller sized for; unused rows must read as zero
        int rows = sgb_spmm_stat_rows(n, c);
        if (rows > grid)
            SGB_CUDA(cudaMemsetAsync(stat_partials + (size_t)grid * 3 * c, 0, (size_t)(rows - grid) * 3 * c * sizeof(float), stream));
    }
    static const int sched = [] { const char* e = getenv("SGB_SPMM_SCHED"); return e ? atoi(e) : 1; }();
    SpmmArgs a{rowptr, reinterpret_cast<const int2*>(edges), dis, mode, x, ldx, n, c, in_mean, in_scale, in_shift, slope, alpha, addend, ld_addend, beta, bias, y, ldy, stat_partials,
               x_ghost, ld_ghost, x_ghost ? (int32_t)n_split : (int32_t)0x7fffffff, sched};
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
    SGB_SPMM_CASE(1, 4, 2)
    SGB_SPMM_CASE(2, 4, 2)
    SGB_SPMM_CASE(4, 4, 2)
    SGB_SPMM_CASE(8, 4, 2)
    SGB_SPMM_CASE(16, 4, 2)
    SGB_SPMM_CASE(32, 4, 2)
    SGB_SPMM_CASE(4, 1, 1)
    SGB_SPMM_CASE(32, 1, 1)
#undef SGB_SPMM_CASE
    set_error("sgb_spmm: no kernel configuration for c=%d", c);
    return SGB_ENOTSUP;
}

// ---------------- row gather (halo pack): out[k, :] = x[idx[k], :] ----------------
namespace sgb {
template <int VEC>
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ idx, int64_t count,
                                                     int c, float* __restrict__ out, int64_t ldo) {
    const int per_row = (c + VEC - 1) / VEC;
    const int64_t total = count * per_row;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / per_row;
        const int ch = (int)(i % per_row) * VEC;
        const int64_t r = idx[k];
        if (VEC == 4) st4(out + k * ldo + ch, ldg4(x + r * ldx + ch));
        else out[k * ldo + ch] = __ldg(x + r * ldx + ch);
    }
}
}  // namespace sgb

extern "C" int sgb_gather_rows(const float* x, int64_t ldx, const int32_t* idx, int64_t count, int c, float* out, int64_t ldo, void* stream_) {
    using namespace sgb;
    cudaStream_t stream = (cudaStream_t)stream_;
    SGB_CHECK_ARG(count >= 0 && c > 0 && ldx >= c && ldo >= c, "sgb_gather_rows: bad shape");
    if (count == 0) return SGB_OK;
    SGB_CHECK_ARG(x && idx && out, "sgb_gather_rows: null pointer");
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = c % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && al16(x) && al16(out);
    const int64_t total = count * (vec ? c / 4 : c);
    const int grid = (int)min64(ceil_div(total, 256), (int64_t)num_sms() * 16);
    if (vec) k_gather_rows<4><<<grid, 256, 0, stream>>>(x, ldx, idx, count, c, out, ldo);
    else k_gather_rows<1><<<grid, 256, 0, stream>>>(x, ldx, idx, count, c, out, ldo);
    SGB_CHECK_LAUNCH("k_gather_rows");
    return SGB_OK;
}
