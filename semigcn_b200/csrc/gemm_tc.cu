// tcgen05 (5th-gen tensor core) tiles for the dense feature transform, fp32-accurate via
// error-compensated 3xTF32:  x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi);
//   A*B ~= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo      (fp32 accumulation in TMEM)
// which keeps ~22 mantissa bits per product -- plain TF32 (10 bits) cannot meet the 1e-5 parity
// bar (SURVEY.md §0 finding 6).
//
// Kernel k_gemm_tc  (forward transform and dX):   C[M,N] (+)= f(A)[M,K] * Wp^T + bias
//   * persistent: one CTA per SM, static round-robin over (m-tile, n-tile) pairs
//   * A (activations, M ~ 1e6 rows) : global -> registers (BatchNorm-affine + LeakyReLU prologue,
//     hi/lo split) -> shared memory in the canonical K-major no-swizzle UMMA layout, padded so the
//     16-byte stores are bank-conflict free                       [8 producer warps]
//   * B (weights, tiny, L2-resident): pre-split and pre-tiled once per call by k_prep_weights into
//     exactly the shared-memory image one stage needs, then fetched with one cp.async.bulk per
//     stage signalling the stage's mbarrier                       [1 elected lane]
//   * MMA: tcgen05.mma.cta_group::1.kind::tf32, M=128, N<=256, K=8 per instruction, 12 per
//     32-wide K chunk, issued by one elected lane; smem slots released with tcgen05.commit
//   * accumulators double-buffered in TMEM (2 x N columns) so the epilogue of tile i overlaps the
//     main loop of tile i+1; epilogue warps read TMEM with tcgen05.ld.32x32b, add bias (and C when
//     accumulating) and store rows straight to global memory      [4 epilogue warps]
//
// Kernel k_gemm_tn_tc (weight gradient):  D[N,K] = G[M,N]^T A[M,K], reduction over the vertices;
//   both operands are MN-major for the tensor core, so the row-major global tiles are staged with
//   16-byte stores as they are (no transposition); split over M across CTAs, fixed-order
//   reduction of the partial tiles afterwards (deterministic).
#include "common.cuh"
#include "umma.cuh"

namespace sgb {

// ------------------------------------------------------------------------------------------
// tile geometry (forward / dX kernel)
// ------------------------------------------------------------------------------------------
constexpr int kTcBM = 128;                      // rows per tile == UMMA M
constexpr int kTcBK = 32;                       // K chunk per pipeline stage (4 MMA k-slices of 8)
constexpr int kTcALbo = 144;                    // bytes between K-adjacent 8x16B core matrices of A (128 + 16 pad: conflict-free stores)
constexpr int kTcASbo = (kTcBK / 4) * kTcALbo;  // bytes between M-adjacent core matrices of A = 1152
constexpr int kTcATile = (kTcBM / 8) * kTcASbo; // 18432 bytes per hi (or lo) A tile
constexpr int kTcBLbo = 128;
constexpr int kTcBSbo = (kTcBK / 4) * kTcBLbo;  // 1024
constexpr int kTcProducerWarps = 8;
constexpr int kTcThreads = (kTcProducerWarps + 2 + 4) * 32;   // producers, B-copy warp, MMA warp, 4 epilogue warps
constexpr int kTcMaxStages = 4;

__host__ __device__ constexpr int tc_b_tile_bytes(int bn) { return (bn / 8) * kTcBSbo; }     // per hi (or lo)
__host__ __device__ constexpr int tc_stage_bytes(int bn) { return 2 * kTcATile + 2 * tc_b_tile_bytes(bn); }

struct TcArgs {
    const float* a; int64_t lda;
    const float* wp;                  // prepped weights: [n_tiles][k_chunks][hi|lo][tile image]
    float* c; int64_t ldc;
    int64_t m; int n; int k;
    int bn;                           // columns per n-tile (multiple of 16, <= 256); all n-tiles but the last are full
    int n_tiles; int k_chunks; int stages;
    const float* a_mean; const float* a_scale; const float* a_shift; float slope;
    const float* bias; int accumulate;
    uint32_t tmem_cols;
    int acc_stride;                   // TMEM columns between the two accumulator buffers (multiple of 32)
};

// weights -> hi/lo split, zero padded, in the exact per-stage shared-memory image
// element (n, k) of n-tile t, chunk q:  core (nl/8, kl/4), row nl%8, word kl%4
__global__ void k_prep_weights(const float* __restrict__ w, int64_t ldw, int transb, int n, int k, int bn, int n_tiles, int k_chunks,
                               float* __restrict__ wp) {
    const int tile_words = tc_b_tile_bytes(bn) / 4;
    const int64_t total = (int64_t)n_tiles * k_chunks * tile_words;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int word = (int)(i % tile_words);
        const int64_t blk = i / tile_words;
        const int q = (int)(blk % k_chunks), t = (int)(blk / k_chunks);
        const int core_n = word / (kTcBSbo / 4);
        const int rem = word % (kTcBSbo / 4);
        const int core_k = rem / 32, in_core = rem % 32;
        const int nl = core_n * 8 + in_core / 4, kl = core_k * 4 + in_core % 4;
        const int gn = t * bn + nl, gk = q * kTcBK + kl;
        float v = 0.f;
        if (gn < n && gk < k) v = transb ? w[(int64_t)gn * ldw + gk] : w[(int64_t)gk * ldw + gn];
        float hi, lo;
        split_tf32(v, hi, lo);
        float* base = wp + (blk * 2) * tile_words;
        base[word] = hi;
        base[tile_words + word] = lo;
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) k_gemm_tc(const TcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stage_bytes = tc_stage_bytes(g.bn);
    const int b_tile_bytes = tc_b_tile_bytes(g.bn);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)g.stages * stage_bytes);
    uint64_t* empty = full + kTcMaxStages;
    uint64_t* tfull = empty + kTcMaxStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(&full[s], kTcProducerWarps * 32 + 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);
        }
        fence_barrier_init();
    }
    if (warp == kTcProducerWarps + 1) tmem_alloc(tmem_slot, g.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t m_tiles = (g.m + kTcBM - 1) / kTcBM;
    const int64_t total_tiles = m_tiles * g.n_tiles;

    if (warp < kTcProducerWarps) {
        // ================= A producers: global -> regs (prologue, hi/lo) -> smem =================
        const int ptid = threadIdx.x;                          // 0..255
        const bool pro = g.a_scale != nullptr;
        // chunk = 128 rows x 8 float4; thread handles rows r0 + 32*i (i<4), float4 column cq
        const int cq = ptid & 7, r0 = ptid >> 3;               // r0 in 0..31
        // flattened (tile, k-chunk) iteration space with a two-deep register prefetch: the loads of
        // iteration i+2 are in flight while iteration i is converted and stored
        const int64_t my_tiles = total_tiles > blockIdx.x ? (total_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
        const int64_t total_it = my_tiles * g.k_chunks;
        auto issue = [&](int64_t i, float4 (&v)[4]) {
            const int64_t tile = blockIdx.x + (i / g.k_chunks) * gridDim.x;
            const int q = (int)(i % g.k_chunks);
            const int64_t m0 = (tile / g.n_tiles) * kTcBM;
            const int kcol = q * kTcBK + cq * 4;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int64_t gm = m0 + r0 + 32 * r;
                v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < total_it && gm < g.m && kcol < g.k) v[r] = ldg4(g.a + gm * g.lda + kcol);   // K % 4 == 0 guaranteed by the dispatcher
            }
        };
        float4 nxt0[4], nxt1[4];      // iterations it and it+1 (plain registers: no dynamic indexing)
        issue(0, nxt0);
        issue(1, nxt1);
        for (int64_t it = 0; it < total_it; ++it) {
            const int s = (int)(it % g.stages);
            const uint32_t ph = (uint32_t)((it / g.stages) & 1);
            const int64_t tile = blockIdx.x + (it / g.k_chunks) * gridDim.x;
            const int q = (int)(it % g.k_chunks);
            const int64_t m0 = (tile / g.n_tiles) * kTcBM;
            const int kcol = q * kTcBK + cq * 4;
            float4 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { v[r] = nxt0[r]; nxt0[r] = nxt1[r]; }
            issue(it + 2, nxt1);
            float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pro && kcol < g.k) { mu = ldg4(g.a_mean + kcol); sc = ldg4(g.a_scale + kcol); sh = ldg4(g.a_shift + kcol); }
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* a_hi = smem + (size_t)s * stage_bytes;
            uint8_t* a_lo = a_hi + kTcATile;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + 32 * i;
                float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
                if (pro) {
                    x[0] = bn_lrelu(x[0], mu.x, sc.x, sh.x, g.slope); x[1] = bn_lrelu(x[1], mu.y, sc.y, sh.y, g.slope);
                    x[2] = bn_lrelu(x[2], mu.z, sc.z, sh.z, g.slope); x[3] = bn_lrelu(x[3], mu.w, sc.w, sh.w, g.slope);
                    if (m0 + r >= g.m || kcol >= g.k) { x[0] = x[1] = x[2] = x[3] = 0.f; }
                }
                float h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_tf32(x[e], h[e], l[e]);
                const uint32_t off = (uint32_t)(r >> 3) * kTcASbo + (uint32_t)cq * kTcALbo + (uint32_t)(r & 7) * 16;
                *reinterpret_cast<float4*>(a_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(a_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&full[s]);
        }
    } else if (warp == kTcProducerWarps) {
        // ================= B copy: one bulk copy of the pre-tiled hi|lo weight image per stage =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = (int)(tile % g.n_tiles);
                for (int q = 0; q < g.k_chunks; ++q, ++it) {
                    const int s = it % g.stages;
                    const uint32_t ph = (it / g.stages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* b_dst = smem + (size_t)s * stage_bytes + 2 * kTcATile;
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(g.wp) + ((size_t)nt * g.k_chunks + q) * 2 * b_tile_bytes;
                    mbar_arrive_expect_tx(&full[s], 2 * b_tile_bytes);
                    bulk_g2s(b_dst, src, 2 * b_tile_bytes, &full[s]);
                }
            }
        }
    } else if (warp == kTcProducerWarps + 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
                const int nt = (int)(tile % g.n_tiles);
                const int ncols = min(g.bn, g.n - nt * g.bn);                 // multiple of 16
                const uint32_t idesc = make_idesc_tf32(kTcBM, ncols, 0, 0);
                const int acc = tcount & 1;
                mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * g.acc_stride);
                for (int q = 0; q < g.k_chunks; ++q, ++it) {
                    const int s = it % g.stages;
                    const uint32_t ph = (it / g.stages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t a_lo = a_hi + kTcATile;
                    const uint32_t b_hi = a_hi + 2 * kTcATile;
                    const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
                    for (int j = 0; j < kTcBK / 8; ++j) {
                        const uint64_t dah = make_desc(a_hi + j * 2 * kTcALbo, kTcALbo, kTcASbo);
                        const uint64_t dal = make_desc(a_lo + j * 2 * kTcALbo, kTcALbo, kTcASbo);
                        const uint64_t dbh = make_desc(b_hi + j * 2 * kTcBLbo, kTcBLbo, kTcBSbo);
                        const uint64_t dbl = make_desc(b_lo + j * 2 * kTcBLbo, kTcBLbo, kTcBSbo);
                        // small terms first, the dominant hi*hi product last
                        umma_tf32(d_tmem, dal, dbh, idesc, (q | j) ? 1u : 0u);
                        umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                        umma_tf32(d_tmem, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&empty[s]);          // frees the smem slot when the MMAs above have read it
                }
                umma_commit(&tfull[acc]);            // accumulator complete
            }
        }
    } else {
        // ================= epilogue: TMEM -> registers -> (+bias, +C) -> global =================
        const int quarter = warp & 3;                // TMEM lanes 32*quarter .. +31 are accessible to this warp
        uint32_t tcount = 0;
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int nt = (int)(tile % g.n_tiles);
            const int64_t m0 = (tile / g.n_tiles) * kTcBM;
            const int n0 = nt * g.bn;
            const int ncols = min(g.bn, g.n - n0);
            const int acc = tcount & 1;
            mbar_wait(&tfull[acc], (tcount >> 1) & 1);
            tc_fence_after();
            const int64_t gm = m0 + quarter * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * g.acc_stride);
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                float v[32];
                tmem_ld_32x32(taddr + c0, v);        // warp-collective: executed by all lanes regardless of row validity
                if (gm < g.m) {
                    float* cp = g.c + gm * g.ldc + n0 + c0;
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        if (c0 + e < ncols) {        // ncols is a multiple of 16 => whole float4 valid
                            float4 o = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                            if (g.accumulate) {
                                const float4 old = *reinterpret_cast<const float4*>(cp + e);
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            if (g.bias) {
                                const float4 b = ldg4(g.bias + n0 + c0 + e);
                                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                            }
                            *reinterpret_cast<float4*>(cp + e) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTcProducerWarps + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, g.tmem_cols);
    }
}

// column moments of C (BatchNorm partials: count, mean, M2) when the 3xTF32 path produced C: CTA b walks the 128-row
// tiles b, b + gridDim.x, ... and Chan-merges them, so the partials buffer holds gridDim.x <= gemm_stat_rows(m) rows
__global__ void __launch_bounds__(256) k_tile_col_stats(const float* __restrict__ c, int64_t ldc, int64_t m, int n, float* __restrict__ partials) {
    const int64_t m_tiles = (m + kTcBM - 1) / kTcBM;
    for (int col = threadIdx.x; col < n; col += blockDim.x) {
        Moments run{0.f, 0.f, 0.f};
        for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
            const int64_t m0 = tile * kTcBM;
            const int rows = (int)min64(kTcBM, m - m0);
            const float p = c[m0 * ldc + col];            // pivot: first row of the tile
            float s1 = 0.f, s2 = 0.f;
            for (int r = 0; r < rows; ++r) {
                const float d = c[(m0 + r) * ldc + col] - p;
                s1 += d;
                s2 = fmaf(d, d, s2);
            }
            run = merge(run, from_shifted((float)rows, p, s1, s2));
        }
        partials[((int64_t)blockIdx.x * 3 + 0) * n + col] = run.n;
        partials[((int64_t)blockIdx.x * 3 + 1) * n + col] = run.mean;
        partials[((int64_t)blockIdx.x * 3 + 2) * n + col] = run.m2;
    }
}

// ------------------------------------------------------------------------------------------
// weight gradient on the tensor pipe:  D[n,k] = sum_m G[m,n] * A[m,k]
//   UMMA M = 128 rows of n, UMMA N = k-tile (<= 256), UMMA K = 8 vertices per instruction.
//   Both operands are MN-major.  For 32-bit types the tensor core accepts MN-major operands only in
//   the SWIZZLE_128B_BASE32B layout (layout type 1; every other layout type reads zeros -- measured
//   with tools/mma_probe.cu, whose output also fixed the address map below):
//     atom = 4 vertices (K) x 128 B (32 consecutive n or k indices), rows 128 B apart, the 32-byte
//     chunk index XOR-ed with the row:  byte(idx, m) = (idx/32)*LBO + (m/4)*SBO + (m%4)*128
//                                                      + ((((idx%32)/8) ^ (m%4)) * 32) + (idx%8)*4
//   so a row-major global row (n or k contiguous) is staged with plain 16-byte stores, no
//   transposition anywhere, and a quarter-warp always fills one whole 128-byte row (conflict-free).
//   One output tile + one m-slice per CTA; slices reduced afterwards in fixed order.
// ------------------------------------------------------------------------------------------
constexpr int kTnBM = 128;                       // rows of D (n) per CTA
constexpr int kTnBK = 32;                        // vertices per pipeline stage
constexpr int kTnSbo = 512;                      // K-direction atom stride (4 vertices x 128 B)
constexpr int kTnLbo = (kTnBK / 4) * kTnSbo;     // MN-direction atom stride = 4096
constexpr int kTnGTile = (kTnBM / 32) * kTnLbo;  // 16384 bytes per hi (or lo)
constexpr uint32_t kLayoutSw128Base32 = 1;
constexpr int kTnProducerWarps = 8;
constexpr int kTnDrainWarps = 4;
constexpr int kTnThreads = (kTnProducerWarps + 1 + kTnDrainWarps) * 32;
// Tensor-core accumulation truncates (measured: error grows ~6e-9 per accumulated vertex row), so one
// TMEM accumulator never sums more than kTnSegChunks * kTnBK = 512 vertices: segments alternate between
// two accumulators and are drained into the CTA's partial tile with round-to-nearest fp32 adds.
constexpr int kTnSegChunks = 16;

__host__ __device__ constexpr int tn_a_tile_bytes(int bk) { return ((bk + 31) / 32) * kTnLbo; }
__host__ __device__ constexpr int tn_stage_bytes(int bk) { return 2 * kTnGTile + 2 * tn_a_tile_bytes(bk); }

struct TnArgs {
    const float* g; int64_t ldg;
    const float* a; int64_t lda;
    float* partial;                   // [splits][n][k]
    int64_t m; int n; int k;
    int bk;                           // UMMA N: columns of D per CTA (multiple of 16, <= 256)
    int k_tiles; int stages;
    int64_t m_per_split;              // multiple of kTnBK
    uint32_t tmem_cols;
    int acc_stride;                   // TMEM columns between the two accumulators (multiple of 32)
};

__global__ void __launch_bounds__(kTnThreads, 1) k_gemm_tn_tc(const TnArgs t) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stage_bytes = tn_stage_bytes(t.bk);
    const int a_tile_bytes = tn_a_tile_bytes(t.bk);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)t.stages * stage_bytes);
    uint64_t* empty = full + kTcMaxStages;
    uint64_t* tfull = empty + kTcMaxStages;      // [2]
    uint64_t* tempty = tfull + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    if (threadIdx.x == 0) {
        for (int s = 0; s < t.stages; ++s) {
            mbar_init(&full[s], kTnProducerWarps * 32);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], kTnDrainWarps);
        }
        fence_barrier_init();
    }
    if (warp == kTnProducerWarps) tmem_alloc(tmem_slot, t.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n0 = (blockIdx.x / t.k_tiles) * kTnBM;
    const int k0 = (blockIdx.x % t.k_tiles) * t.bk;
    const int64_t ms = (int64_t)blockIdx.y * t.m_per_split;
    const int64_t me = min64(t.m, ms + t.m_per_split);
    const int chunks = me > ms ? (int)((me - ms + kTnBK - 1) / kTnBK) : 0;
    const int segments = (chunks + kTnSegChunks - 1) / kTnSegChunks;

    if (warp < kTnProducerWarps) {
        // each warp owns 4 of the 32 vertex rows of a stage; a lane owns 16-byte pieces lane, lane+32
        const int a_pieces = t.bk / 4;             // <= 64
        for (int it = 0; it < chunks; ++it) {
            const int s = it % t.stages;
            const uint32_t ph = (it / t.stages) & 1;
            float4 gv[4], av[4][2];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int64_t gm = ms + (int64_t)it * kTnBK + warp * 4 + r;
                const bool rok = gm < me;
                const int gn = n0 + lane * 4;
                gv[r] = (rok && gn < t.n) ? ldg4(t.g + gm * t.ldg + gn) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int piece = lane + 32 * h;
                    const int gk = k0 + piece * 4;
                    av[r][h] = (rok && piece < a_pieces && gk < t.k) ? ldg4(t.a + gm * t.lda + gk) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* g_hi = smem + (size_t)s * stage_bytes;
            uint8_t* g_lo = g_hi + kTnGTile;
            uint8_t* a_hi = g_lo + kTnGTile;
            uint8_t* a_lo = a_hi + a_tile_bytes;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int ml = warp * 4 + r;                                   // vertex row within the stage
                const uint32_t row_off = (uint32_t)(ml >> 2) * kTnSbo + (uint32_t)(ml & 3) * 128;
                const uint32_t rx = (uint32_t)(ml & 3);
                {
                    float x[4] = {gv[r].x, gv[r].y, gv[r].z, gv[r].w}, h4[4], l4[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_tf32(x[e], h4[e], l4[e]);
                    const uint32_t off = (uint32_t)(lane >> 3) * kTnLbo + row_off + ((((uint32_t)(lane & 7) >> 1) ^ rx) << 5) + (uint32_t)(lane & 1) * 16;
                    *reinterpret_cast<float4*>(g_hi + off) = make_float4(h4[0], h4[1], h4[2], h4[3]);
                    *reinterpret_cast<float4*>(g_lo + off) = make_float4(l4[0], l4[1], l4[2], l4[3]);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int piece = lane + 32 * h;
                    if (piece < a_pieces) {
                        float x[4] = {av[r][h].x, av[r][h].y, av[r][h].z, av[r][h].w}, h4[4], l4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) split_tf32(x[e], h4[e], l4[e]);
                        const uint32_t off = (uint32_t)(piece >> 3) * kTnLbo + row_off + ((((uint32_t)(piece & 7) >> 1) ^ rx) << 5) + (uint32_t)(piece & 1) * 16;
                        *reinterpret_cast<float4*>(a_hi + off) = make_float4(h4[0], h4[1], h4[2], h4[3]);
                        *reinterpret_cast<float4*>(a_lo + off) = make_float4(l4[0], l4[1], l4[2], l4[3]);
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
        }
    } else if (warp == kTnProducerWarps) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(kTnBM, t.bk, 1, 1);      // both operands MN-major
            for (int it = 0; it < chunks; ++it) {
                const int s = it % t.stages;
                const uint32_t ph = (it / t.stages) & 1;
                const int seg = it / kTnSegChunks, in_seg = it % kTnSegChunks;
                const int acc = seg & 1;
                if (in_seg == 0) {
                    mbar_wait(&tempty[acc], ((seg >> 1) & 1) ^ 1);          // drained (passes at once for the first two segments)
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * t.acc_stride);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t g_hi = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t g_lo = g_hi + kTnGTile;
                const uint32_t a_hi = g_lo + kTnGTile;
                const uint32_t a_lo = a_hi + a_tile_bytes;
#pragma unroll
                for (int j = 0; j < kTnBK / 8; ++j) {
                    const uint64_t dgh = make_desc(g_hi + j * 2 * kTnSbo, kTnLbo, kTnSbo, kLayoutSw128Base32);
                    const uint64_t dgl = make_desc(g_lo + j * 2 * kTnSbo, kTnLbo, kTnSbo, kLayoutSw128Base32);
                    const uint64_t dah = make_desc(a_hi + j * 2 * kTnSbo, kTnLbo, kTnSbo, kLayoutSw128Base32);
                    const uint64_t dal = make_desc(a_lo + j * 2 * kTnSbo, kTnLbo, kTnSbo, kLayoutSw128Base32);
                    umma_tf32(d_tmem, dgl, dah, idesc, (in_seg | j) ? 1u : 0u);
                    umma_tf32(d_tmem, dgh, dal, idesc, 1u);
                    umma_tf32(d_tmem, dgh, dah, idesc, 1u);
                }
                umma_commit(&empty[s]);
                if (in_seg == kTnSegChunks - 1 || it == chunks - 1) umma_commit(&tfull[acc]);
            }
        }
    } else {
        // drain warps: TMEM lane quarter = warp % 4; running total lives in the CTA's partial tile (L2-resident),
        // always updated by the same thread in the same order -> deterministic
        const int quarter = warp & 3;
        float* out = t.partial + (int64_t)blockIdx.y * t.n * t.k;
        const int gn = n0 + quarter * 32 + lane;
        const int kcols = min(t.bk, t.k - k0);
        const bool vec_ok = (t.k % 4 == 0) && (k0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        if (segments == 0) {
            if (gn < t.n)
                for (int c = 0; c < kcols; ++c) out[(int64_t)gn * t.k + k0 + c] = 0.f;
        }
        for (int seg = 0; seg < segments; ++seg) {
            const int acc = seg & 1;
            mbar_wait(&tfull[acc], (seg >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * t.acc_stride);
            for (int c0 = 0; c0 < kcols; c0 += 32) {
                float v[32];
                tmem_ld_32x32(taddr + (uint32_t)c0, v);
                if (gn < t.n) {
                    float* op = out + (int64_t)gn * t.k + k0 + c0;
                    if (vec_ok && c0 + 32 <= kcols) {
                        float4 old[8];
                        if (seg != 0) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) old[e] = *reinterpret_cast<const float4*>(op + 4 * e);
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float4 o = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                            if (seg != 0) { o.x += old[e].x; o.y += old[e].y; o.z += old[e].z; o.w += old[e].w; }
                            *reinterpret_cast<float4*>(op + 4 * e) = o;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (c0 + e < kcols) op[e] = seg == 0 ? v[e] : op[e] + v[e];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTnProducerWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, t.tmem_cols);
    }
}

__global__ void k_reduce_splits_tc(const float* __restrict__ partial, int splits, int n, int k, float* __restrict__ d, int64_t ldd,
                                   int accumulate) {
    const int64_t total = (int64_t)n * k;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sidx = 0; sidx < splits; ++sidx) s += partial[(int64_t)sidx * total + i];
        float* dp = d + (i / k) * ldd + (i % k);
        *dp = accumulate ? *dp + s : s;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int tile_col_stats_launch(const float* c, int64_t ldc, int64_t m, int n, float* partials, cudaStream_t stream) {
    const int rows = gemm_stat_rows(m);
    const int grid = (int)min64(ceil_div(m, kTcBM), rows);
    if (rows > grid) SGB_CUDA(cudaMemsetAsync(partials + (size_t)grid * 3 * n, 0, (size_t)(rows - grid) * 3 * n * sizeof(float), stream));
    k_tile_col_stats<<<grid, 256, 0, stream>>>(c, ldc, m, n, partials);
    SGB_CHECK_LAUNCH("k_tile_col_stats");
    return SGB_OK;
}

int reduce_splits_launch(const float* partial, int splits, int n, int k, float* d, int64_t ldd, int accumulate, cudaStream_t stream) {
    int64_t total = (int64_t)n * k;
    int rgrid = (int)min64(ceil_div(total, 256), (int64_t)num_sms() * 8);
    k_reduce_splits_tc<<<rgrid, 256, 0, stream>>>(partial, splits, n, k, d, ldd, accumulate);
    SGB_CHECK_LAUNCH("k_reduce_splits");
    return SGB_OK;
}

struct TcPlan {
    int bn, n_tiles, k_chunks, stages, acc_stride;
    size_t smem_bytes, ws_bytes;
    uint32_t tmem_cols;
};

static const int kSmemBudget = 227 * 1024;

static TcPlan tc_plan(int n, int k) {
    TcPlan p;
    p.bn = n <= 256 ? n : 256;
    p.n_tiles = (n + p.bn - 1) / p.bn;
    p.k_chunks = (k + kTcBK - 1) / kTcBK;
    int sb = tc_stage_bytes(p.bn);
    int st = (kSmemBudget - 256) / sb;
    p.stages = st > kTcMaxStages ? kTcMaxStages : st;
    p.smem_bytes = (size_t)p.stages * sb + 256;
    p.ws_bytes = (size_t)p.n_tiles * p.k_chunks * 2 * tc_b_tile_bytes(p.bn);
    p.acc_stride = (p.bn + 31) / 32 * 32;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * p.acc_stride)) cols <<= 1;
    p.tmem_cols = cols;
    return p;
}

bool gemm_tc_supported(int64_t m, int n, int k, int64_t lda, int64_t ldc, const void* a, const void* c, const void* bias,
                       const void* a_scale) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return n >= 16 && n % 16 == 0 && k >= 8 && k % 4 == 0 && lda % 4 == 0 && ldc % 4 == 0 && al16(a) && al16(c) &&
           (!bias || al16(bias)) && (!a_scale || al16(a_scale)) && m >= 1;
}

size_t gemm_tc_workspace(int n, int k) { return tc_plan(n, k).ws_bytes + 256; }

int gemm_tc_launch(const GemmArgs& g, void* ws, size_t ws_bytes, cudaStream_t stream) {
    TcPlan p = tc_plan(g.n, g.k);
    if (!ws || ws_bytes < p.ws_bytes) {
        set_error("sgb_gemm: tensor-core engine needs %zu bytes of workspace, got %zu", p.ws_bytes, ws_bytes);
        return SGB_ENOSPC;
    }
    if (p.stages < 2) {
        set_error("sgb_gemm: tile does not fit shared memory");
        return SGB_ENOTSUP;
    }
    float* wp = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    {
        int64_t total = (int64_t)p.n_tiles * p.k_chunks * (tc_b_tile_bytes(p.bn) / 4);
        int grid = (int)min64(ceil_div(total, 256), (int64_t)num_sms() * 8);
        k_prep_weights<<<grid, 256, 0, stream>>>(g.b, g.ldb, g.transb, g.n, g.k, p.bn, p.n_tiles, p.k_chunks, wp);
        SGB_CHECK_LAUNCH("k_prep_weights");
    }
    {
        static std::atomic<uint64_t> optin{0};
        int rc = smem_optin(reinterpret_cast<const void*>(k_gemm_tc), kSmemBudget, &optin);
        if (rc != SGB_OK) return rc;
    }
    TcArgs t{};
    t.a = g.a; t.lda = g.lda; t.wp = wp; t.c = g.c; t.ldc = g.ldc; t.m = g.m; t.n = g.n; t.k = g.k;
    t.bn = p.bn; t.n_tiles = p.n_tiles; t.k_chunks = p.k_chunks; t.stages = p.stages;
    t.a_mean = g.a_mean; t.a_scale = g.a_scale; t.a_shift = g.a_shift; t.slope = g.slope; t.bias = g.bias; t.accumulate = g.accumulate;
    t.tmem_cols = p.tmem_cols; t.acc_stride = p.acc_stride;
    int64_t tiles = ceil_div(g.m, kTcBM) * p.n_tiles;
    int grid = (int)min64(tiles, num_sms());
    k_gemm_tc<<<grid, kTcThreads, p.smem_bytes, stream>>>(t);
    SGB_CHECK_LAUNCH("k_gemm_tc");
    if (g.stat_partials) return tile_col_stats_launch(g.c, g.ldc, g.m, g.n, g.stat_partials, stream);
    return SGB_OK;
}

// ---- weight gradient ----
struct TnPlan {
    int bk, k_tiles, n_tiles, stages, splits, acc_stride;
    int64_t m_per_split;
    size_t smem_bytes, ws_bytes;
    uint32_t tmem_cols;
};

static TnPlan tn_plan(int64_t m, int n, int k) {
    TnPlan p;
    int kpad = (k + 15) / 16 * 16;
    p.bk = kpad <= 256 ? kpad : 256;
    p.k_tiles = (k + p.bk - 1) / p.bk;
    p.n_tiles = (n + kTnBM - 1) / kTnBM;
    int sb = tn_stage_bytes(p.bk);
    int st = (kSmemBudget - 256) / sb;
    p.stages = st > kTcMaxStages ? kTcMaxStages : st;
    p.smem_bytes = (size_t)p.stages * sb + 256;
    int tiles = p.k_tiles * p.n_tiles;
    int64_t want = num_sms() / tiles;
    if (want < 1) want = 1;
    int64_t maxs = ceil_div(m > 0 ? m : 1, 1024);
    p.splits = (int)(want < maxs ? want : maxs);
    p.m_per_split = ceil_div(ceil_div(m, p.splits), kTnBK) * kTnBK;
    p.ws_bytes = (size_t)p.splits * n * k * sizeof(float);
    p.acc_stride = (p.bk + 31) / 32 * 32;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * p.acc_stride)) cols <<= 1;
    p.tmem_cols = cols;
    return p;
}

bool gemm_tn_tc_supported(int64_t m, int n, int k, int64_t ldg, int64_t lda, const void* g, const void* a) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return n % 4 == 0 && k % 4 == 0 && ldg % 4 == 0 && lda % 4 == 0 && al16(g) && al16(a) && m >= 1;
}

size_t gemm_tn_tc_workspace(int64_t m, int n, int k) { return tn_plan(m, n, k).ws_bytes; }

int gemm_tn_tc_launch(const float* g, int64_t ldg, const float* a, int64_t lda, float* d, int64_t ldd, int64_t m, int n, int k,
                      int accumulate, void* ws, size_t ws_bytes, cudaStream_t stream) {
    TnPlan p = tn_plan(m, n, k);
    if (!ws || ws_bytes < p.ws_bytes) {
        set_error("sgb_gemm_tn: tensor-core engine needs %zu bytes of workspace, got %zu", p.ws_bytes, ws_bytes);
        return SGB_ENOSPC;
    }
    if (p.stages < 2) {
        set_error("sgb_gemm_tn: tile does not fit shared memory");
        return SGB_ENOTSUP;
    }
    {
        static std::atomic<uint64_t> optin{0};
        int rc = smem_optin(reinterpret_cast<const void*>(k_gemm_tn_tc), kSmemBudget, &optin);
        if (rc != SGB_OK) return rc;
    }
    TnArgs t{};
    t.g = g; t.ldg = ldg; t.a = a; t.lda = lda; t.partial = (float*)ws; t.m = m; t.n = n; t.k = k;
    t.bk = p.bk; t.k_tiles = p.k_tiles; t.stages = p.stages; t.m_per_split = p.m_per_split; t.tmem_cols = p.tmem_cols; t.acc_stride = p.acc_stride;
    dim3 grid((unsigned)(p.n_tiles * p.k_tiles), (unsigned)p.splits);
    k_gemm_tn_tc<<<grid, kTnThreads, p.smem_bytes, stream>>>(t);
    SGB_CHECK_LAUNCH("k_gemm_tn_tc");
    int64_t total = (int64_t)n * k;
    int rgrid = (int)min64(ceil_div(total, 256), (int64_t)num_sms() * 8);
    k_reduce_splits_tc<<<rgrid, 256, 0, stream>>>((const float*)ws, p.splits, n, k, d, ldd, accumulate);
    SGB_CHECK_LAUNCH("k_reduce_splits");
    return SGB_OK;
}

}  // namespace sgb
