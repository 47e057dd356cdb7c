// Persistent, warp-specialised bf16 GEMM on the sm_100a tensor cores.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : MMA issuer     (one lane issues tcgen05.mma, accumulators in TMEM, 2 buffers)
//   warps 2..9  : epilogue       (tcgen05.ld -> registers -> fused epilogue -> global), 2 warps per TMEM quadrant
//
// The same kernel serves
//   * linear layers and their gradients (K-major and MN-major operands, split-K with fp32 red),
//   * 3x3 / 1x1 convolution forward and data-gradient as implicit GEMM: the A tile of a
//     k-iteration is one TMA box {64 channels, bw, bh} of the NHWC activation, shifted by the
//     filter tap; out-of-image pixels are zero-filled by the TMA unit (that is the padding),
//   * convolution weight-gradient (reduction over pixels, both operands MN-major, taps folded into N),
//   * the batched attention GEMMs with softmax-aware epilogues.
//
// Replaces in the reference: nn.Linear / nn.Conv2d dispatch sites K1,K3,K4,K5 of SURVEY.md §2.2
// (modules/diffusion/openaimodel.py:247-301, modules/attention.py:50-74,283-290,616-639).
#include "gemm_tc.cuh"

#include <algorithm>
#include <cstdlib>

namespace nk {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int ATOM_BYTES = 64 * 64 * 2;     // one 64x64 bf16 box, 8 KB
constexpr int NUM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;

// MODE_CONV_FWD_T: convolution with <= 128 output channels computed transposed (D = W x pixels^T): the 128 output
// channels are the UMMA M dimension and 256 PIXELS the N dimension, because a UMMA costs the same ~150 clk for N = 128
// and N = 256 (measured), so pixel-major tiles with N = Cout = 128 run the tensor pipe at half rate.
enum { MODE_PLAIN = 0, MODE_CONV_FWD = 1, MODE_CONV_WGRAD = 2, MODE_CONV_FWD_T = 3 };

// barrier of one epilogue warp-half (4 warps) with a compile-time barrier id: a register id makes ptxas reserve all 16
// hardware barriers, and such a kernel fails to launch ("too many resources requested")
__device__ __forceinline__ void half_bar_sync(int half) {
    if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
}

constexpr int STG_BYTES = 128 * 128;  // one staging tile of the TMA-store epilogue: 128 rows x 128 bytes, 128B-swizzled

struct alignas(64) GemmDev {
    CUtensorMap tmA;
    CUtensorMap tmB;
    CUtensorMap tmC;   // output tile store (tma_out): bf16 boxes {64 cols, rows} / fp32 boxes {32 cols, rows}
    CUtensorMap tmC2;  // EPI_GEGLU_FWD: the gated output [M, D]; with epi_prefetch: the epilogue's side input (residual / GEGLU h)
    int tma_out;       // 1: the epilogue stages 128-byte-wide column groups in shared memory and stores them with TMA
    int stg_per_half;  // staging tiles per epilogue warp-half
    int has_c;         // EPI_GEGLU_FWD: h is stored as well
    int M, N;
    int BN;
    int tiles_m, tiles_n, nb2, nb1, splits;  // tiles_m counts (CG*128)-row tiles
    int tiles_m128;                          // 128-row tiles actually present (validity of a CTA's half)
    int k_iters, k_iters_total;
    int stages;
    int a_mn, b_mn;
    int mode;
    int dbg_skip;                // NK_GEMM_DBG_SKIP bit 1: no A loads, bit 2: no B loads (timing experiments, wrong results)
    int raster_gm;               // > 0: tiles are walked in groups of raster_gm row blocks x all N tiles (see launch_gemm)
    int dual_skew;               // DUAL: k-iterations by which row tile 1 trails row tile 0 (clamped to stages - 1 in the kernel)
    int epi_prefetch;            // 1: tmC2 maps the epilogue's side input (residual, or h of EPI_GEGLU_BWD); its boxes are prefetched into L2
    int a_b2, a_b1, b_b2, b_b1;  // 0/1: does the operand carry that batch dimension
    // conv geometry
    int cH, cW, bw, bh, tiles_w, tiles_h, cin_blocks, ksize, pad;  // cH, cW: OUTPUT image (= input unless strided)
    int cstride, pad_t, pad_l;                                     // conv forward: input pixel = cstride*out + tap - pad
    // epilogue
    void* C;
    long long ldc, c_b2, c_b1;
    int out, epi;
    float alpha;
    const float* bias;
    const float* bias_img;
    int rows_per_img;
    const bf16* residual;
    long long ldr;
    const float* rowvec;
    const bf16* aux;
    int geglu_d;       // EPI_GEGLU_*: D (the tile scheduler runs over N = D columns)
    void* C2;          // EPI_GEGLU_FWD: gated output [M, D]
    long long ldc2;
    uint32_t idesc;
    uint32_t a_bytes, b_bytes;  // bytes landed per stage for A and B
    long long* dbg;             // optional per-CTA cycle counters (NK_GEMM_DEBUG_TIMING), 8 per CTA
};

struct TileCoord {
    int mt, nt, b2, b1, split;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmDev& g, int t) {
    TileCoord c;
    if (g.raster_gm > 0) {
        const int per_batch = g.tiles_m * g.tiles_n;
        const int tb = t % per_batch;
        t /= per_batch;
        const int per_group = g.raster_gm * g.tiles_n;
        const int grp = tb / per_group;
        const int r = tb - grp * per_group;
        const int m0 = grp * g.raster_gm;
        const int gsize = min(g.raster_gm, g.tiles_m - m0);  // the last group may be short
        c.nt = r / gsize;
        c.mt = m0 + (r - c.nt * gsize);
    } else {
        c.mt = t % g.tiles_m;
        t /= g.tiles_m;
        c.nt = t % g.tiles_n;
        t /= g.tiles_n;
    }
    c.b2 = t % g.nb2;
    t /= g.nb2;
    c.b1 = t % g.nb1;
    c.split = t / g.nb1;
    return c;
}

__device__ __forceinline__ int iters_of_split(const GemmDev& g, int split) {
    int rem = g.k_iters_total - split * g.k_iters;
    return rem < g.k_iters ? rem : g.k_iters;
}

// v[0..15] += p[0..15] (fp32 vector, same address for every lane of the warp -> broadcast loads)
__device__ __forceinline__ void add_vec16(float (&v)[16], const float* p, bool full, int remaining) {
    if (full && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(p) + j);
            v[4 * j] += f.x;
            v[4 * j + 1] += f.y;
            v[4 * j + 2] += f.z;
            v[4 * j + 3] += f.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < remaining) v[j] += __ldg(p + j);
    }
}

// VAR selects an epilogue family that lives in its own instantiation (so that its register pressure cannot spill into the
// main kernel): 0 = general, 1 = MODE_CONV_FWD_T (transposed convolution), 2 = EPI_GEGLU_FWD, 3 = EPI_GEGLU_BWD
enum { VAR_MAIN = 0, VAR_TRANSPOSED = 1, VAR_GEGLU_FWD = 2, VAR_GEGLU_BWD = 3 };
// TMA_OUT: the epilogue leaves through shared-memory staging tiles and TMA stores (g.tma_out says the same at run time)
// DUAL (CTA pairs, VAR_MAIN): one scheduler tile is TWO row tiles that share the B tile — 256 rows per CTA, 512 per pair.
//   The kernel is bound by the L2 -> SM operand feed (profiles/r02_gemm_vs_cublas.txt: cuBLAS and this kernel both move
//   operands at ~9.1 TB/s, cuBLAS needs 25 % fewer bytes because its CTA owns 256 x 256 outputs); a stage then holds
//   A0 | A1 | B-half = 48 KB for twice the MMA work of the 32 KB single-tile stage.  The two row tiles accumulate in the
//   two 256-column TMEM buffers at the same time (no accumulator double buffering across tiles); the epilogue warps see
//   them as two ordinary 128-row tiles (accumulator = sub-tile), so every epilogue family is unchanged.  While the
//   epilogue drains, the producer keeps filling the ring, which is what matters in the feed-bound regime.
template <int CG, int VAR = VAR_MAIN, bool TMA_OUT = false, bool DUAL = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmDev g) {
    constexpr bool TRANSPOSED = VAR == VAR_TRANSPOSED;
    static_assert(!DUAL || (CG == 2 && VAR == VAR_MAIN), "DUAL: CTA pairs, general epilogue family only");
    constexpr int A_STG = DUAL ? 2 * A_STAGE_BYTES : A_STAGE_BYTES;  // bytes of one ring slot's A part
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024B alignment is required by the 128B swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    const int stages = g.stages;
    const int BNc = g.BN / CG;  // rows of the B tile held by this CTA
    const uint32_t b_stage_bytes = static_cast<uint32_t>(BNc) * 128u;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int num_groups = gridDim.x / CG;
    const int group = blockIdx.x / CG;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + static_cast<size_t>(stages) * A_STG;
    uint8_t* smem_stg = smem_b + static_cast<size_t>(stages) * b_stage_bytes;  // 1024-aligned (stage sizes are multiples of 2 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + (TMA_OUT ? 2 * g.stg_per_half * STG_BYTES : 0));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + stages;
    uint64_t* tmem_full = bars + 2 * stages;
    uint64_t* tmem_empty = bars + 2 * stages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = g.tiles_m * g.tiles_n * g.nb2 * g.nb1 * g.splits;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&g.tmA);
        tma_prefetch_desc(&g.tmB);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full[0], 1);
        mbar_init(&tmem_full[1], 1);
        mbar_init(&tmem_empty[0], 8 * CG);
        mbar_init(&tmem_empty[1], 8 * CG);
        fence_barrier_init();
    }
    if (CG == 2) cluster_sync_all();  // barrier inits visible to the peer CTA before anything is signalled
    if (warp == 1) {
        if (CG == 2) {
            tmem_alloc_2cta(tmem_slot, TMEM_COLS);
            tmem_relinquish_2cta();
        } else {
            tmem_alloc(tmem_slot, TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp walks the k-loop; lane 0 waits for the ring slot and posts the expected byte count, then
        // lane l issues the l-th TMA load of the k-iteration (up to 6: MN-major operands are loaded per 64-wide atom),
        // so the issue latency of several cp.async.bulk.tensor instructions is not serialised on one thread.
        {
            int stage = 0;
            uint32_t phase = 0;
            long long dbg_acc0 = 0;
            auto tma_load = [](const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
                if (CG == 2) tma_load_4d_2cta(tm, bar, dst, c0, c1, c2, c3);
                else tma_load_4d(tm, bar, dst, c0, c1, c2, c3);
            };
            const int nb_atoms = BNc / 64;  // MN-major / wgrad B tiles: 64-wide atoms per CTA
            for (int t = group; t < total_tiles; t += num_groups) {
                const TileCoord tc = decode_tile(g, t);
                const int n0 = tc.nt * g.BN + static_cast<int>(rank) * BNc;  // this CTA's slice of the B tile
                // this CTA's 128-row tile (DUAL: the first of its two; the second lies CG tiles = one pair tile further)
                const int mt = (DUAL ? 2 * tc.mt : tc.mt) * CG + static_cast<int>(rank);
                const int m0 = mt * BM;
                const int iters = iters_of_split(g, tc.split);
                int img = 0, th = 0, tw = 0;
                int img1 = 0, th1 = 0, tw1 = 0;  // DUAL, conv forward: pixel tile of the second row tile
                // Every integer division in this k-loop is ~40 clk of latency that delays all operand traffic of the CTA
                // (the conv weight gradient spent 1 500 clk per k-iteration here, 3x the tensor time).  Everything that
                // depends only on the tile is computed once; the k-dependent indices (filter tap / channel block, or the
                // pixel tile) advance incrementally.
                int cb = 0, dy = 0, dx = 0;      // conv forward: k-iteration -> (tap = (dy, dx), 64-channel block)
                int ptw = 0, pth = 0, pimg = 0;  // conv wgrad:   k-iteration -> 64-pixel tile
                int my_c0 = 0, my_dx = 0, my_dy = 0;  // conv wgrad: channel block and tap shift of THIS lane's B atom
                const int gi0 = tc.split * g.k_iters;
                if (g.mode == MODE_CONV_FWD || g.mode == MODE_CONV_FWD_T) {
                    const int ptile = (g.mode == MODE_CONV_FWD_T) ? tc.nt : mt;  // pixel tile of this CTA
                    tw = ptile % g.tiles_w;
                    th = (ptile / g.tiles_w) % g.tiles_h;
                    img = ptile / (g.tiles_w * g.tiles_h);  // may be >= nimg for the odd last tile: TMA zero-fills
                    if (DUAL) {
                        const int ptile1 = mt + CG;
                        tw1 = ptile1 % g.tiles_w;
                        th1 = (ptile1 / g.tiles_w) % g.tiles_h;
                        img1 = ptile1 / (g.tiles_w * g.tiles_h);
                    }
                    const int tap = gi0 / g.cin_blocks;
                    cb = gi0 - tap * g.cin_blocks;
                    dy = tap / g.ksize;
                    dx = tap - dy * g.ksize;
                } else if (g.mode == MODE_CONV_WGRAD) {
                    ptw = gi0 % g.tiles_w;
                    pth = (gi0 / g.tiles_w) % g.tiles_h;
                    pimg = gi0 / (g.tiles_w * g.tiles_h);
                    // every 64-channel atom of the B tile carries its own filter tap (its own pixel shift), so a
                    // 256-wide N tile may straddle taps; atoms past N are fetched out of range (zero-filled)
                    const int atom = (n0 >> 6) + max(lane - 2, 0);
                    const int t_ = atom / g.cin_blocks;
                    const bool in_range = t_ < g.ksize * g.ksize;
                    my_c0 = in_range ? ((atom - t_ * g.cin_blocks) << 6) : (g.cin_blocks << 6);
                    my_dy = t_ / g.ksize - g.pad;
                    my_dx = t_ % g.ksize - g.pad;
                }
                const int x0 = g.cstride * (tw * g.bw) - g.pad_l, y0 = g.cstride * (th * g.bh) - g.pad_t;
                const int x1 = g.cstride * (tw1 * g.bw) - g.pad_l, y1 = g.cstride * (th1 * g.bh) - g.pad_t;
                for (int it = 0; it < iters; ++it) {
                    const int gi = gi0 + it;
                    if (lane == 0) {
                        const long long tw0 = g.dbg ? clock64() : 0;
                        mbar_wait(&empty_bar[stage], phase ^ 1u, 100u + stage);
                        if (g.dbg) dbg_acc0 += clock64() - tw0;
                        if (rank == 0)
                            mbar_arrive_expect_tx(&full_bar[stage], CG * (((g.dbg_skip & 1) ? 0u : (DUAL ? 2u : 1u) * g.a_bytes) +
                                                                          ((g.dbg_skip & 2) ? 0u : g.b_bytes)));
                    }
                    __syncwarp();
                    uint8_t* sa = smem_a + static_cast<size_t>(stage) * A_STG;
                    uint8_t* sb = smem_b + static_cast<size_t>(stage) * b_stage_bytes;
                    if (VAR == VAR_GEGLU_FWD) {
                        // B = W [2D, K]: the tile's first BN/2 rows are value features [nt*BN/2, ...), the last BN/2
                        // rows the gate rows of the same features (D rows further down).  A CTA pair gets that split for
                        // free — rank 0 holds the value half of the UMMA's N range, rank 1 the gate half.
                        const int k0 = gi * BK;
                        const int hr = g.BN / 2;
                        if (lane == 0) {
                            tma_load(&g.tmA, &full_bar[stage], sa, k0, m0, 0, 0);
                        } else if (CG == 2) {
                            if (lane == 1)
                                tma_load(&g.tmB, &full_bar[stage], sb, k0, static_cast<int>(rank) * g.geglu_d + tc.nt * hr, 0, 0);
                        } else if (lane <= 2) {
                            const int j = lane - 1;
                            tma_load(&g.tmB, &full_bar[stage], sb + static_cast<size_t>(j) * hr * 128, k0,
                                     j * g.geglu_d + tc.nt * hr, 0, 0);
                        }
                    } else if (g.mode == MODE_PLAIN) {
                        const int k0 = gi * BK;
                        const int na = (g.a_mn ? 2 : 1) * (DUAL ? 2 : 1);
                        const int nb = g.b_mn ? nb_atoms : 1;
                        if (g.dbg_skip && ((lane < na && (g.dbg_skip & 1)) || (lane >= na && (g.dbg_skip & 2)))) {
                            // timing experiment: this operand is not fetched (stale shared memory is multiplied)
                        } else if (lane < na) {
                            // DUAL: lanes [0, na/2) fetch the first row tile, [na/2, na) the second (A1 follows A0 in the slot)
                            const int sub = DUAL ? (g.a_mn ? (lane >> 1) : lane) : 0;
                            const int at = g.a_mn ? (lane & 1) : 0;  // 64-row atom of an MN-major A tile
                            const int ms = m0 + sub * (CG * BM);
                            uint8_t* dst = sa + sub * A_STAGE_BYTES + at * ATOM_BYTES;
                            if (!g.a_mn)
                                tma_load(&g.tmA, &full_bar[stage], dst, k0, ms, tc.b2 * g.a_b2, tc.b1 * g.a_b1);
                            else
                                tma_load(&g.tmA, &full_bar[stage], dst, ms + 64 * at, k0, tc.b2 * g.a_b2, tc.b1 * g.a_b1);
                        } else if (lane < na + nb) {
                            const int j = lane - na;
                            if (!g.b_mn)
                                tma_load(&g.tmB, &full_bar[stage], sb, k0, n0, tc.b2 * g.b_b2, tc.b1 * g.b_b1);
                            else
                                tma_load(&g.tmB, &full_bar[stage], sb + j * ATOM_BYTES, n0 + 64 * j, k0, tc.b2 * g.b_b2,
                                         tc.b1 * g.b_b1);
                        }
                    } else if (g.mode == MODE_CONV_FWD || g.mode == MODE_CONV_FWD_T) {
                        if (g.mode == MODE_CONV_FWD) {
                            if (lane == 0) tma_load(&g.tmA, &full_bar[stage], sa, cb * 64, x0 + dx, y0 + dy, img);
                            else if (lane == 1) tma_load(&g.tmB, &full_bar[stage], sb, gi * BK, n0, 0, 0);
                            else if (DUAL && lane == 2)
                                tma_load(&g.tmA, &full_bar[stage], sa + A_STAGE_BYTES, cb * 64, x1 + dx, y1 + dy, img1);
                        } else {  // transposed: A = packed weights [Cout, taps*Cin], B = 256-pixel box
                            if (lane == 0) tma_load(&g.tmA, &full_bar[stage], sa, gi * BK, m0, 0, 0);
                            else if (lane == 1) tma_load(&g.tmB, &full_bar[stage], sb, cb * 64, x0 + dx, y0 + dy, img);
                        }
                        if (++cb == g.cin_blocks) {  // next filter tap
                            cb = 0;
                            if (++dx == g.ksize) {
                                dx = 0;
                                ++dy;
                            }
                        }
                    } else {  // MODE_CONV_WGRAD: k-iteration = one 64-pixel tile; N index = tap*Cin + ci
                        const int px = ptw * g.bw, py = pth * g.bh;
                        if (lane < 2)
                            tma_load(&g.tmA, &full_bar[stage], sa + lane * ATOM_BYTES, m0 + 64 * lane, px, py, pimg);
                        else if (lane < 2 + nb_atoms)
                            tma_load(&g.tmB, &full_bar[stage], sb + (lane - 2) * ATOM_BYTES, my_c0, px + my_dx, py + my_dy,
                                     pimg);
                        if (++ptw == g.tiles_w) {  // next pixel tile
                            ptw = 0;
                            if (++pth == g.tiles_h) {
                                pth = 0;
                                ++pimg;
                            }
                        }
                    }
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
            if (g.dbg && lane == 0) g.dbg[blockIdx.x * 8 + 0] = dbg_acc0;
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA of the pair only) =====================
        if (lane == 0 && rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            long long dbg_full = 0, dbg_tempty = 0;
            const long long dbg_t0 = g.dbg ? clock64() : 0;
            // per-k16 descriptor advance (in 16-byte units)
            const uint32_t a_adv = g.a_mn ? (2048u >> 4) : (32u >> 4);
            const uint32_t b_adv = g.b_mn ? (2048u >> 4) : (32u >> 4);
            const uint32_t a_lbo = g.a_mn ? static_cast<uint32_t>(ATOM_BYTES) : 16u;
            const uint32_t b_lbo = g.b_mn ? static_cast<uint32_t>(ATOM_BYTES) : 16u;
            const uint64_t a_desc0 = make_smem_desc(smem_u32(smem_a), a_lbo, 1024u);
            const uint64_t b_desc0 = make_smem_desc(smem_u32(smem_b), b_lbo, 1024u);
            if constexpr (DUAL) {
                // Both accumulators belong to the current scheduler tile: buffer s = row tile s, same B slot for both.
                // Row tile 1 trails row tile 0 by `sk` k-iterations (g.dual_skew, at most stages - 1: the slot of
                // k-iteration j is reused by j + stages, and it is row tile 1 that frees it).  The lead lets row tile 0
                // start while the epilogue still drains buffer 1 of the previous scheduler tile, and hands buffer 0 to
                // the epilogue `sk` k-iterations before buffer 1 is complete; sk = 0 is the plain interleaved order.
                for (int t = group; t < total_tiles; t += num_groups, ++local) {
                    const TileCoord tc = decode_tile(g, t);
                    const int iters = iters_of_split(g, tc.split);
                    const int sk = min(min(g.dual_skew, stages - 1), iters);
                    const uint32_t par = (static_cast<uint32_t>(local) & 1u) ^ 1u;  // n-th use of each buffer, n = local
                    int st1 = stage;  // ring position of row tile 1 (row tile 0 walks `stage` / `phase`)
                    for (int j = 0; j < iters + sk; ++j) {
                        if (j < iters) {  // ---- row tile 0, k-iteration j
                            long long tw = g.dbg ? clock64() : 0;
                            mbar_wait(&full_bar[stage], phase, 300u + stage);
                            if (g.dbg) dbg_full += clock64() - tw;
                            tc_fence_after();
                            if (j == 0) {  // the epilogue of the previous scheduler tile has drained buffer 0
                                tw = g.dbg ? clock64() : 0;
                                mbar_wait(&tmem_empty[0], par, 200u);
                                if (g.dbg) dbg_tempty += clock64() - tw;
                                tc_fence_after();
                            }
                            const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((stage * A_STG) >> 4);
                            const uint64_t b_desc = b_desc0 + static_cast<uint64_t>((stage * b_stage_bytes) >> 4);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                tc_mma_ss_2cta(tmem_base, a_desc + static_cast<uint64_t>(k * a_adv),
                                               b_desc + static_cast<uint64_t>(k * b_adv), g.idesc, (j > 0 || k > 0) ? 1u : 0u);
                            if (j == iters - 1) tc_commit_2cta(&tmem_full[0]);  // row tile 0 complete -> epilogue
                            if (++stage == stages) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                        if (j >= sk) {  // ---- row tile 1, k-iteration j - sk (its slot was waited for by row tile 0)
                            const int j1 = j - sk;
                            if (j1 == 0) {
                                const long long tw = g.dbg ? clock64() : 0;
                                mbar_wait(&tmem_empty[1], par, 201u);
                                if (g.dbg) dbg_tempty += clock64() - tw;
                                tc_fence_after();
                            }
                            const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((st1 * A_STG + A_STAGE_BYTES) >> 4);
                            const uint64_t b_desc = b_desc0 + static_cast<uint64_t>((st1 * b_stage_bytes) >> 4);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                tc_mma_ss_2cta(tmem_base + static_cast<uint32_t>(ACC_STRIDE),
                                               a_desc + static_cast<uint64_t>(k * a_adv),
                                               b_desc + static_cast<uint64_t>(k * b_adv), g.idesc, (j1 > 0 || k > 0) ? 1u : 0u);
                            if (j1 == iters - 1) tc_commit_2cta(&tmem_full[1]);  // row tile 1 complete -> epilogue
                            tc_commit_2cta(&empty_bar[st1]);  // frees the smem slot: both row tiles have read it
                            if (++st1 == stages) st1 = 0;
                        }
                    }
                }
            } else
            for (int t = group; t < total_tiles; t += num_groups, ++local) {
                const TileCoord tc = decode_tile(g, t);
                const int iters = iters_of_split(g, tc.split);
                const int acc = local & 1;
                long long tw = g.dbg ? clock64() : 0;
                mbar_wait(&tmem_empty[acc], ((local >> 1) & 1u) ^ 1u, 200u + acc);
                if (g.dbg) dbg_tempty += clock64() - tw;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * ACC_STRIDE);
                for (int it = 0; it < iters; ++it) {
                    tw = g.dbg ? clock64() : 0;
                    mbar_wait(&full_bar[stage], phase, 300u + stage);
                    if (g.dbg) dbg_full += clock64() - tw;
                    tc_fence_after();
                    const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((stage * A_STAGE_BYTES) >> 4);
                    const uint64_t b_desc = b_desc0 + static_cast<uint64_t>((stage * b_stage_bytes) >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if (CG == 2)
                            tc_mma_ss_2cta(d_tmem, a_desc + static_cast<uint64_t>(k * a_adv),
                                           b_desc + static_cast<uint64_t>(k * b_adv), g.idesc,
                                           (it > 0 || k > 0) ? 1u : 0u);
                        else
                            tc_mma_ss(d_tmem, a_desc + static_cast<uint64_t>(k * a_adv),
                                      b_desc + static_cast<uint64_t>(k * b_adv), g.idesc,
                                      (it > 0 || k > 0) ? 1u : 0u);
                    }
                    if (CG == 2) tc_commit_2cta(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (CG == 2) tc_commit_2cta(&tmem_full[acc]); else tc_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
            }
            if (g.dbg) {
                g.dbg[blockIdx.x * 8 + 1] = dbg_full;
                g.dbg[blockIdx.x * 8 + 2] = dbg_tempty;
                g.dbg[blockIdx.x * 8 + 3] = clock64() - dbg_t0;
                g.dbg[blockIdx.x * 8 + 6] = local;
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // Two warps per TMEM lane quadrant, each owning half of the tile's 16-column chunks.  TMEM reads and the
        // side input (residual / aux) of chunk c+1 are issued before chunk c is processed, so their latency overlaps.
        const int quad = warp & 3;          // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;   // which half of the column chunks
        const int r = quad * 32 + lane;
        const int nch = g.BN / 16;
        const int c_begin = half ? (nch + 1) / 2 : 0;
        const int c_end = half ? nch : (nch + 1) / 2;
        int local = 0;
        int stg_it = 0;  // staging-buffer alternation of the TMA-store epilogue (runs across tiles)
        const bool st_elected = (((warp - 2) & 3) == 0) && lane == 0;  // the thread of this warp-half that issues TMA stores
        long long dbg_wait = 0, dbg_proc = 0;
        for (;; ++local) {
            // DUAL: scheduler tile local >> 1, row tile (= accumulator) local & 1 — from here on an ordinary 128-row tile
            const int t = group + (DUAL ? (local >> 1) : local) * num_groups;
            if (t >= total_tiles) break;
            TileCoord tc = decode_tile(g, t);
            if (DUAL) tc.mt = 2 * tc.mt + (local & 1);
            const int n0 = tc.nt * g.BN;
            const int mt = tc.mt * CG + static_cast<int>(rank);
            const int acc = local & 1;
            // coordinates of this CTA's output tile in the store tensor map (dims 1..3)
            int sc1, sc2, sc3;
            if (g.mode == MODE_CONV_FWD) {
                sc1 = (mt % g.tiles_w) * g.bw;
                sc2 = ((mt / g.tiles_w) % g.tiles_h) * g.bh;
                sc3 = mt / (g.tiles_w * g.tiles_h);
            } else {
                sc1 = mt * BM;
                sc2 = tc.b2;
                sc3 = tc.b1;
            }
            const bool tile_ok = mt < g.tiles_m128;  // the odd last 128-row tile of a CTA pair stores nothing
            // output row of this thread
            long long row = 0;
            bool row_ok = false;
            int img = 0;
            if (g.mode == MODE_CONV_FWD) {
                const int tw = mt % g.tiles_w;
                const int th = (mt / g.tiles_w) % g.tiles_h;
                img = mt / (g.tiles_w * g.tiles_h);
                const int hh = r / g.bw;
                const int ww = r - hh * g.bw;
                const int h = th * g.bh + hh;
                const int w = tw * g.bw + ww;
                row_ok = (mt < g.tiles_m128) && (hh < g.bh) && (h < g.cH) && (w < g.cW);
                row = (static_cast<long long>(img) * g.cH + h) * g.cW + w;
            } else {
                row = static_cast<long long>(mt) * BM + r;
                row_ok = row < g.M;
                if (g.bias_img) img = static_cast<int>(row / g.rows_per_img);
            }
            const long long boff = tc.b2 * g.c_b2 + tc.b1 * g.c_b1;
            float rv = 0.f;
            if (g.rowvec && row_ok)
                rv = g.rowvec[(static_cast<long long>(tc.b1) * g.nb2 + tc.b2) * g.M + row];
            // side input row pointer: residual (EPI_LINEAR) or aux (EPI_DSOFTMAX)
            const bf16* side_row = nullptr;
            if (row_ok) {
                if (g.epi == EPI_LINEAR && g.residual) side_row = g.residual + boff + row * g.ldr;
                else if (g.epi == EPI_DSOFTMAX) side_row = g.aux + boff + row * g.ldc;
            }
            const float* bias_img_row = g.bias_img ? g.bias_img + static_cast<long long>(img) * g.N : nullptr;

            if (TMA_OUT && g.epi_prefetch && st_elected && tile_ok) {
                // The side input of this warp-half's column groups goes to L2 NOW, while the main loop of this tile still
                // runs: the epilogue reads it with one 32-byte load per thread = row and chunk (32 rows per warp
                // instruction), which is latency-bound when the rows come from DRAM (the GEGLU data gradient reads
                // 335 MB of h per call this way and runs at half the rate of the other GEMMs).  A hint only.
                if constexpr (VAR == VAR_GEGLU_BWD) {
                    const int ngp = g.BN / 64;
                    const int pb = half ? (ngp + 1) / 2 : 0, pe = half ? ngp : (ngp + 1) / 2;
#pragma unroll 1
                    for (int gi = pb; gi < pe; ++gi) {
                        const int col = n0 + gi * 64;
                        if (col < g.geglu_d) {
                            tma_prefetch_l2_4d(&g.tmC2, col, sc1, sc2, sc3);
                            tma_prefetch_l2_4d(&g.tmC2, g.geglu_d + col, sc1, sc2, sc3);
                        }
                    }
                } else if constexpr (VAR == VAR_MAIN) {
                    const int ngp = (g.BN + 63) / 64;
                    const int pb = half ? (ngp + 1) / 2 : 0, pe = half ? ngp : (ngp + 1) / 2;
#pragma unroll 1
                    for (int gi = pb; gi < pe; ++gi) {
                        const int col = n0 + gi * 64;
                        if (col < g.N) tma_prefetch_l2_4d(&g.tmC2, col, sc1, sc2, sc3);
                    }
                }
            }
            const long long te0 = g.dbg ? clock64() : 0;
            mbar_wait(&tmem_full[acc], (local >> 1) & 1u, 400u + acc);
            const long long te1 = g.dbg ? clock64() : 0;
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                    static_cast<uint32_t>(acc * ACC_STRIDE);

            auto side_vec_ok = [&](int c) -> bool {
                const int n = n0 + c * 16;
                return side_row != nullptr && (n + 16 <= g.N) &&
                       ((reinterpret_cast<uintptr_t>(side_row + n) & 15u) == 0);
            };
            auto issue = [&](int c, uint32_t (&raw)[16], uint4& sa, uint4& sb) {
                tc_ld16(taddr0 + static_cast<uint32_t>(c * 16), raw);
                if (side_vec_ok(c)) {
                    const uint4* sp = reinterpret_cast<const uint4*>(side_row + n0 + c * 16);
                    sa = __ldg(sp);
                    sb = __ldg(sp + 1);
                }
            };
            // stg != nullptr: the 16 results go to row r of a 128B-swizzled staging tile (chunk `cig` of its column group)
            // and leave through a TMA store; rows / columns outside the problem are clipped by the TMA unit
            auto process = [&](int c, const uint32_t (&raw)[16], const uint4& sa, const uint4& sb, uint8_t* stg = nullptr,
                               int cig = 0) {
                const int n = n0 + c * 16;
                if (!row_ok || n >= g.N) return;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
                const bool full = (n + 16 <= g.N);
                // side values (fp32) when present
                float sv[16];
                const bool has_side = side_row != nullptr;
                if (has_side) {
                    if (side_vec_ok(c)) {
                        const uint32_t w[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 f = unpack_bf16x2(w[j]);
                            sv[2 * j] = f.x;
                            sv[2 * j + 1] = f.y;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) sv[j] = (n + j < g.N) ? __bfloat162float(side_row[n + j]) : 0.f;
                    }
                }
                if (g.epi == EPI_LINEAR) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] *= g.alpha;
                    if (g.bias) add_vec16(v, g.bias + n, full, g.N - n);
                    if (bias_img_row) add_vec16(v, bias_img_row + n, full, g.N - n);
                    if (has_side) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += sv[j];
                    }
                } else if (g.epi == EPI_EXP2) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = exp2f(v[j] * g.alpha - rv);
                } else {  // EPI_DSOFTMAX
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = sv[j] * (v[j] - rv) * g.alpha;
                }
                // ---- store ----
                if (TMA_OUT) {
                    const uint32_t rowaddr = smem_u32(stg) + static_cast<uint32_t>(r * 128);
                    const uint32_t sw = static_cast<uint32_t>(r & 7);
                    if (g.out == OUT_BF16) {
                        const uint32_t j0 = static_cast<uint32_t>(2 * cig);
                        st_shared_v4(rowaddr + ((j0 ^ sw) << 4),
                                     make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                                pack_bf16x2(v[6], v[7])));
                        st_shared_v4(rowaddr + (((j0 + 1) ^ sw) << 4),
                                     make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                                pack_bf16x2(v[14], v[15])));
                    } else {
                        const uint32_t j0 = static_cast<uint32_t>(4 * cig);
#pragma unroll
                        for (uint32_t j = 0; j < 4; ++j)
                            st_shared_v4(rowaddr + (((j0 + j) ^ sw) << 4),
                                         make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                    __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
                    }
                    return;
                }
                if (g.out == OUT_BF16) {
                    bf16* cp = reinterpret_cast<bf16*>(g.C) + boff + row * g.ldc + n;
                    if (full && ((reinterpret_cast<uintptr_t>(cp) & 15u) == 0)) {
                        uint4 q0, q1;
                        q0.x = pack_bf16x2(v[0], v[1]);
                        q0.y = pack_bf16x2(v[2], v[3]);
                        q0.z = pack_bf16x2(v[4], v[5]);
                        q0.w = pack_bf16x2(v[6], v[7]);
                        q1.x = pack_bf16x2(v[8], v[9]);
                        q1.y = pack_bf16x2(v[10], v[11]);
                        q1.z = pack_bf16x2(v[12], v[13]);
                        q1.w = pack_bf16x2(v[14], v[15]);
                        reinterpret_cast<uint4*>(cp)[0] = q0;
                        reinterpret_cast<uint4*>(cp)[1] = q1;
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n + j < g.N) cp[j] = __float2bfloat16(v[j]);
                    }
                } else if (g.out == OUT_F32) {
                    float* cp = reinterpret_cast<float*>(g.C) + boff + row * g.ldc + n;
                    if (full && ((reinterpret_cast<uintptr_t>(cp) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            reinterpret_cast<float4*>(cp)[j] =
                                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n + j < g.N) cp[j] = v[j];
                    }
                } else {  // OUT_F32_ATOMIC
                    float* cp = reinterpret_cast<float*>(g.C) + boff + row * g.ldc + n;
                    if (full && ((reinterpret_cast<uintptr_t>(cp) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * j),
                                         "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                         : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n + j < g.N) atomicAdd(cp + j, v[j]);
                    }
                }
            };
            if constexpr (TRANSPOSED) {
                // D[co (TMEM lane), pixel (column)]: thread = one output channel, chunk = 16 consecutive pixels of the
                // tile.  For every pixel the 32 lanes of a warp write 32 consecutive channels (64 contiguous bytes).
                const int pt = tc.nt;
                const int ttw = pt % g.tiles_w, tth = (pt / g.tiles_w) % g.tiles_h, timg = pt / (g.tiles_w * g.tiles_h);
                const int co = tc.mt * BM + r;
                const bool co_ok = co < g.N;
                float bco = 0.f;
                if (co_ok && g.bias) bco += g.bias[co];
                if (co_ok && g.bias_img) bco += g.bias_img[static_cast<long long>(timg) * g.N + co];
                const int npx = g.bw * g.bh;
                bf16* cbase = reinterpret_cast<bf16*>(g.C) + co;
                const int co_c = co_ok ? co : 0;  // clamped: loads below are unconditional (their results are masked)
                const unsigned short* rbase_c =
                    g.residual ? reinterpret_cast<const unsigned short*>(g.residual) + co_c : nullptr;
                // Two chunks in flight: the TMEM load and the 16 independent 2-byte residual loads of chunk c+1 are issued
                // before chunk c is converted and stored.  (A load under a per-element branch is waited for at the branch
                // join — 440 clk per pixel, 56 k clk per tile in the first version — so loads are unconditional on
                // clamped addresses and masked afterwards.)
                struct ChunkT {
                    long long base;   // pixel index (image-global) of the chunk's first column, clamped into the tensor
                    int nv;           // 16: all columns valid and consecutive in one image row (fast path); else -1
                    int off[16];      // slow path: per-column pixel offsets inside the image (clamped)
                    uint32_t okmask;  // slow path: validity of each column
                };
                const long long img_base = static_cast<long long>(timg) * g.cH * g.cW;
                const bool row_chunks = (g.bw & 15) == 0;  // a 16-pixel chunk never straddles two image rows
                auto issue_t = [&](int c, uint32_t (&raw)[16], unsigned short (&rs)[16], ChunkT& ck) {
                    tc_ld16(taddr0 + static_cast<uint32_t>(c * 16), raw);
                    const int p0 = c * 16;
                    const int hh0 = p0 / g.bw, ww0 = p0 - hh0 * g.bw;
                    const int h0 = tth * g.bh + hh0, w0 = ttw * g.bw + ww0;
                    if (row_chunks && co_ok && hh0 < g.bh && h0 < g.cH && w0 + 16 <= g.cW) {
                        ck.nv = 16;
                        ck.base = img_base + static_cast<long long>(h0) * g.cW + w0;
                        if (rbase_c) {
                            const unsigned short* rp = rbase_c + ck.base * g.ldr;
#pragma unroll
                            for (int j = 0; j < 16; ++j) rs[j] = __ldg(rp + j * g.ldr);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) rs[j] = 0;
                        }
                        return;
                    }
                    ck.nv = -1;
                    int hh = hh0, ww = ww0;
                    ck.okmask = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int h = tth * g.bh + hh, w = ttw * g.bw + ww;
                        const bool ok = co_ok && (p0 + j < npx) && (h < g.cH) && (w < g.cW);
                        ck.okmask |= ok ? (1u << j) : 0u;
                        ck.off[j] = min(h, g.cH - 1) * g.cW + min(w, g.cW - 1);
                        if (++ww == g.bw) {
                            ww = 0;
                            ++hh;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        rs[j] = rbase_c ? __ldg(rbase_c + (img_base + ck.off[j]) * g.ldr) : static_cast<unsigned short>(0);
                };
                auto process_t = [&](const uint32_t (&raw)[16], const unsigned short (&rs)[16], const ChunkT& ck) {
                    if (ck.nv == 16) {
                        bf16* cp = cbase + ck.base * g.ldc;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float rj = __uint_as_float(static_cast<uint32_t>(rs[j]) << 16);  // bf16 -> fp32
                            cp[j * g.ldc] = __float2bfloat16(fmaf(__uint_as_float(raw[j]), g.alpha, bco + rj));
                        }
                        return;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float rj = __uint_as_float(static_cast<uint32_t>(rs[j]) << 16);
                        if ((ck.okmask >> j) & 1u)
                            cbase[(img_base + ck.off[j]) * g.ldc] =
                                __float2bfloat16(fmaf(__uint_as_float(raw[j]), g.alpha, bco + rj));
                    }
                };
                if (c_begin < c_end) {
                    uint32_t ra[16], rb[16];
                    unsigned short sa_[16], sb_[16];
                    ChunkT ca, cb_;
                    issue_t(c_begin, ra, sa_, ca);
                    for (int c = c_begin; c < c_end; c += 2) {
                        tc_wait_ld16(ra);
                        if (c + 1 < c_end) issue_t(c + 1, rb, sb_, cb_);
                        process_t(ra, sa_, ca);
                        if (c + 1 < c_end) {
                            tc_wait_ld16(rb);
                            if (c + 2 < c_end) issue_t(c + 2, ra, sa_, ca);
                            process_t(rb, sb_, cb_);
                        }
                    }
                }
            } else if constexpr (VAR == VAR_GEGLU_FWD) {
                // value chunk c (columns [16c, 16c+16) of the tile) pairs with gate chunk nx + c; the two warps of a lane
                // quadrant split the value chunks.  h = acc + bias is rounded to bf16 FIRST (it is what the backward
                // reads), the gate uses the rounded values.
                const int D = g.geglu_d;
                const int nx = nch / 2;
                const int cb = half ? (nx + 1) / 2 : 0, ce = half ? nx : (nx + 1) / 2;
                const int n0x = tc.nt * (g.BN / 2);
                bf16* hrow = g.C ? reinterpret_cast<bf16*>(g.C) + row * g.ldc : nullptr;
                bf16* orow = reinterpret_cast<bf16*>(g.C2) + row * g.ldc2;
                auto ld2 = [&](int c, uint32_t (&rx)[16], uint32_t (&rg)[16]) {
                    tc_ld16(taddr0 + static_cast<uint32_t>(c * 16), rx);
                    tc_ld16(taddr0 + static_cast<uint32_t>((nx + c) * 16), rg);
                };
                auto proc = [&](int c, const uint32_t (&rx)[16], const uint32_t (&rg)[16]) {
                    const int n = n0x + c * 16;
                    if (!row_ok || n >= D) return;
                    float v[16], t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        v[j] = __uint_as_float(rx[j]);
                        t[j] = __uint_as_float(rg[j]);
                    }
                    if (g.bias) {
                        add_vec16(v, g.bias + n, true, 16);
                        add_vec16(t, g.bias + D + n, true, 16);
                    }
                    uint32_t hv[8], hg[8], ov[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        hv[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                        hg[j] = pack_bf16x2(t[2 * j], t[2 * j + 1]);
                        const float2 a = unpack_bf16x2(hv[j]), b = unpack_bf16x2(hg[j]);
                        ov[j] = pack_bf16x2(a.x * gelu_erf(b.x), a.y * gelu_erf(b.y));
                    }
                    if (hrow) {
                        reinterpret_cast<uint4*>(hrow + n)[0] = make_uint4(hv[0], hv[1], hv[2], hv[3]);
                        reinterpret_cast<uint4*>(hrow + n)[1] = make_uint4(hv[4], hv[5], hv[6], hv[7]);
                        reinterpret_cast<uint4*>(hrow + D + n)[0] = make_uint4(hg[0], hg[1], hg[2], hg[3]);
                        reinterpret_cast<uint4*>(hrow + D + n)[1] = make_uint4(hg[4], hg[5], hg[6], hg[7]);
                    }
                    reinterpret_cast<uint4*>(orow + n)[0] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
                    reinterpret_cast<uint4*>(orow + n)[1] = make_uint4(ov[4], ov[5], ov[6], ov[7]);
                };
                if constexpr (TMA_OUT) {
                    // value column groups of 64: three staging tiles per group (h value part, h gate part, gated output)
                    const int ng = nx / 4;
                    const int gb = half ? (ng + 1) / 2 : 0, ge = half ? ng : (ng + 1) / 2;
                    uint8_t* t_hv = smem_stg + static_cast<size_t>(half * 3) * STG_BYTES;
                    uint8_t* t_hg = t_hv + STG_BYTES;
                    uint8_t* t_o = t_hg + STG_BYTES;
                    const uint32_t rowoff = static_cast<uint32_t>(r * 128), sw = static_cast<uint32_t>(r & 7);
                    for (int gi = gb; gi < ge; ++gi) {
                        uint32_t ax[16], ag[16], bx[16], bg[16];
                        ld2(gi * 4, ax, ag);
                        if (st_elected) tma_wait_group_read<0>();
                        half_bar_sync(half);
                        auto emit = [&](int cig, const uint32_t (&rx)[16], const uint32_t (&rg)[16]) {
                            const int n = n0x + (gi * 4 + cig) * 16;
                            if (!row_ok || n >= D) return;
                            float v[16], tt[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                v[j] = __uint_as_float(rx[j]);
                                tt[j] = __uint_as_float(rg[j]);
                            }
                            if (g.bias) {
                                add_vec16(v, g.bias + n, true, 16);
                                add_vec16(tt, g.bias + D + n, true, 16);
                            }
                            uint32_t hv[8], hg[8], ov[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                hv[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                                hg[j] = pack_bf16x2(tt[2 * j], tt[2 * j + 1]);
                                const float2 a = unpack_bf16x2(hv[j]), b = unpack_bf16x2(hg[j]);
                                ov[j] = pack_bf16x2(a.x * gelu_erf(b.x), a.y * gelu_erf(b.y));
                            }
                            const uint32_t o0 = rowoff + (((2u * cig) ^ sw) << 4), o1 = rowoff + (((2u * cig + 1u) ^ sw) << 4);
                            if (g.has_c) {
                                st_shared_v4(smem_u32(t_hv) + o0, make_uint4(hv[0], hv[1], hv[2], hv[3]));
                                st_shared_v4(smem_u32(t_hv) + o1, make_uint4(hv[4], hv[5], hv[6], hv[7]));
                                st_shared_v4(smem_u32(t_hg) + o0, make_uint4(hg[0], hg[1], hg[2], hg[3]));
                                st_shared_v4(smem_u32(t_hg) + o1, make_uint4(hg[4], hg[5], hg[6], hg[7]));
                            }
                            st_shared_v4(smem_u32(t_o) + o0, make_uint4(ov[0], ov[1], ov[2], ov[3]));
                            st_shared_v4(smem_u32(t_o) + o1, make_uint4(ov[4], ov[5], ov[6], ov[7]));
                        };
                        tc_wait_ld16(ax);
                        tc_wait_ld16(ag);
                        ld2(gi * 4 + 1, bx, bg);
                        emit(0, ax, ag);
                        tc_wait_ld16(bx);
                        tc_wait_ld16(bg);
                        ld2(gi * 4 + 2, ax, ag);
                        emit(1, bx, bg);
                        tc_wait_ld16(ax);
                        tc_wait_ld16(ag);
                        ld2(gi * 4 + 3, bx, bg);
                        emit(2, ax, ag);
                        tc_wait_ld16(bx);
                        tc_wait_ld16(bg);
                        emit(3, bx, bg);
                        fence_proxy_async_smem();
                        half_bar_sync(half);
                        if (st_elected) {
                            const int col = n0x + gi * 64;
                            if (tile_ok && col < D) {
                                if (g.has_c) {
                                    tma_store_4d(&g.tmC, t_hv, col, sc1, sc2, sc3);
                                    tma_store_4d(&g.tmC, t_hg, D + col, sc1, sc2, sc3);
                                }
                                tma_store_4d(&g.tmC2, t_o, col, sc1, sc2, sc3);
                            }
                            tma_commit_group();
                        }
                    }
                } else if (cb < ce) {
                    uint32_t ax[16], ag[16], bx[16], bg[16];
                    ld2(cb, ax, ag);
                    for (int c = cb; c < ce; c += 2) {
                        tc_wait_ld16(ax);
                        tc_wait_ld16(ag);
                        if (c + 1 < ce) ld2(c + 1, bx, bg);
                        proc(c, ax, ag);
                        if (c + 1 < ce) {
                            tc_wait_ld16(bx);
                            tc_wait_ld16(bg);
                            if (c + 2 < ce) ld2(c + 2, ax, ag);
                            proc(c + 1, bx, bg);
                        }
                    }
                }
            } else if constexpr (VAR == VAR_GEGLU_BWD) {
                // acc = d(out)[row, n..n+16); h = (value | gate) rows of the saved pre-activation (aux, row stride ldr):
                //   dh[:, n] = acc * gelu(gate)      dh[:, D + n] = acc * value * gelu'(gate)
                const int D = g.geglu_d;
                const bf16* hrow = row_ok ? g.aux + row * g.ldr : nullptr;
                bf16* drow = reinterpret_cast<bf16*>(g.C) + row * g.ldc;
                auto issue_b = [&](int c, uint32_t (&raw)[16], uint4 (&hx)[2], uint4 (&hg)[2]) {
                    tc_ld16(taddr0 + static_cast<uint32_t>(c * 16), raw);
                    const int n = n0 + c * 16;
                    if (hrow != nullptr && n < D) {
                        const uint4* px = reinterpret_cast<const uint4*>(hrow + n);
                        const uint4* pg = reinterpret_cast<const uint4*>(hrow + D + n);
                        hx[0] = __ldg(px);
                        hx[1] = __ldg(px + 1);
                        hg[0] = __ldg(pg);
                        hg[1] = __ldg(pg + 1);
                    }
                };
                auto proc_b = [&](int c, const uint32_t (&raw)[16], const uint4 (&hx)[2], const uint4 (&hg)[2],
                                  uint8_t* stg = nullptr, int cig = 0) {
                    const int n = n0 + c * 16;
                    if (!row_ok || n >= D) return;
                    const uint32_t wx[8] = {hx[0].x, hx[0].y, hx[0].z, hx[0].w, hx[1].x, hx[1].y, hx[1].z, hx[1].w};
                    const uint32_t wg[8] = {hg[0].x, hg[0].y, hg[0].z, hg[0].w, hg[1].x, hg[1].y, hg[1].z, hg[1].w};
                    uint32_t da[8], dg[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 a = unpack_bf16x2(wx[j]), b = unpack_bf16x2(wg[j]);
                        const float d0 = __uint_as_float(raw[2 * j]) * g.alpha, d1 = __uint_as_float(raw[2 * j + 1]) * g.alpha;
                        da[j] = pack_bf16x2(d0 * gelu_erf(b.x), d1 * gelu_erf(b.y));
                        dg[j] = pack_bf16x2(d0 * a.x * dgelu_erf(b.x), d1 * a.y * dgelu_erf(b.y));
                    }
                    if (TMA_OUT) {  // tile 0: value-part gradient, tile 1: gate-part gradient
                        const uint32_t rowoff = smem_u32(stg) + static_cast<uint32_t>(r * 128), sw = static_cast<uint32_t>(r & 7);
                        const uint32_t o0 = rowoff + (((2u * cig) ^ sw) << 4), o1 = rowoff + (((2u * cig + 1u) ^ sw) << 4);
                        st_shared_v4(o0, make_uint4(da[0], da[1], da[2], da[3]));
                        st_shared_v4(o1, make_uint4(da[4], da[5], da[6], da[7]));
                        st_shared_v4(o0 + STG_BYTES, make_uint4(dg[0], dg[1], dg[2], dg[3]));
                        st_shared_v4(o1 + STG_BYTES, make_uint4(dg[4], dg[5], dg[6], dg[7]));
                        return;
                    }
                    reinterpret_cast<uint4*>(drow + n)[0] = make_uint4(da[0], da[1], da[2], da[3]);
                    reinterpret_cast<uint4*>(drow + n)[1] = make_uint4(da[4], da[5], da[6], da[7]);
                    reinterpret_cast<uint4*>(drow + D + n)[0] = make_uint4(dg[0], dg[1], dg[2], dg[3]);
                    reinterpret_cast<uint4*>(drow + D + n)[1] = make_uint4(dg[4], dg[5], dg[6], dg[7]);
                };
                if constexpr (TMA_OUT) {
                    const int ng = g.BN / 64;
                    const int gb = half ? (ng + 1) / 2 : 0, ge = half ? ng : (ng + 1) / 2;
                    uint8_t* buf = smem_stg + static_cast<size_t>(half * 2) * STG_BYTES;
                    for (int gi = gb; gi < ge; ++gi) {
                        uint32_t r0[16], r1[16];
                        uint4 x0[2], g0[2], x1[2], g1[2];
                        issue_b(gi * 4, r0, x0, g0);
                        if (st_elected) tma_wait_group_read<0>();
                        half_bar_sync(half);
#pragma unroll
                        for (int cc = 0; cc < 4; cc += 2) {
                            tc_wait_ld16(r0);
                            issue_b(gi * 4 + cc + 1, r1, x1, g1);
                            proc_b(gi * 4 + cc, r0, x0, g0, buf, cc);
                            tc_wait_ld16(r1);
                            if (cc + 2 < 4) issue_b(gi * 4 + cc + 2, r0, x0, g0);
                            proc_b(gi * 4 + cc + 1, r1, x1, g1, buf, cc + 1);
                        }
                        fence_proxy_async_smem();
                        half_bar_sync(half);
                        if (st_elected) {
                            const int col = n0 + gi * 64;
                            if (tile_ok && col < D) {
                                tma_store_4d(&g.tmC, buf, col, sc1, sc2, sc3);
                                tma_store_4d(&g.tmC, buf + STG_BYTES, D + col, sc1, sc2, sc3);
                            }
                            tma_commit_group();
                        }
                    }
                } else if (c_begin < c_end) {
                    uint32_t r0[16], r1[16];
                    uint4 x0[2], g0[2], x1[2], g1[2];
                    issue_b(c_begin, r0, x0, g0);
                    for (int c = c_begin; c < c_end; c += 2) {
                        tc_wait_ld16(r0);
                        if (c + 1 < c_end) issue_b(c + 1, r1, x1, g1);
                        proc_b(c, r0, x0, g0);
                        if (c + 1 < c_end) {
                            tc_wait_ld16(r1);
                            if (c + 2 < c_end) issue_b(c + 2, r0, x0, g0);
                            proc_b(c + 1, r1, x1, g1);
                        }
                    }
                }
            } else if constexpr (TMA_OUT) {
                // 128-byte-wide column groups (64 bf16 / 32 fp32 columns): TMEM -> registers -> swizzled staging tile ->
                // ONE TMA store (or reduce-add) per group, double buffered per warp-half.  Replaces per-thread 32-byte
                // global stores whose 32 lanes hit 32 different rows (32 LSU wavefronts per instruction).
                const int gw = (g.out == OUT_BF16) ? 64 : 32;
                const int cpg = gw / 16;
                const int ng = g.BN / gw;
                const int gb = half ? (ng + 1) / 2 : 0, ge = half ? ng : (ng + 1) / 2;
                for (int gi = gb; gi < ge; ++gi, ++stg_it) {
                    uint8_t* buf = smem_stg + static_cast<size_t>(half * 2 + (stg_it & 1)) * STG_BYTES;
                    const int c0 = gi * cpg;
                    uint32_t r0[16], r1[16];
                    uint4 s0a = make_uint4(0, 0, 0, 0), s0b = s0a, s1a = s0a, s1b = s0a;
                    issue(c0, r0, s0a, s0b);
                    if (st_elected) tma_wait_group_read<1>();  // the store issued two groups ago has released this buffer
                    half_bar_sync(half);
                    for (int c = c0; c < c0 + cpg; c += 2) {
                        tc_wait_ld16(r0);
                        issue(c + 1, r1, s1a, s1b);
                        process(c, r0, s0a, s0b, buf, c - c0);
                        tc_wait_ld16(r1);
                        if (c + 2 < c0 + cpg) issue(c + 2, r0, s0a, s0b);
                        process(c + 1, r1, s1a, s1b, buf, c + 1 - c0);
                    }
                    fence_proxy_async_smem();
                    half_bar_sync(half);
                    if (st_elected) {
                        const int col = n0 + gi * gw;
                        if (tile_ok && col < g.N) {
                            if (g.out == OUT_F32_ATOMIC) tma_reduce_add_4d(&g.tmC, buf, col, sc1, sc2, sc3);
                            else tma_store_4d(&g.tmC, buf, col, sc1, sc2, sc3);
                        }
                        tma_commit_group();
                    }
                }
            } else if (c_begin < c_end) {
                uint32_t r0[16], r1[16];
                uint4 s0a = make_uint4(0, 0, 0, 0), s0b = s0a, s1a = s0a, s1b = s0a;
                issue(c_begin, r0, s0a, s0b);
                for (int c = c_begin; c < c_end; c += 2) {
                    tc_wait_ld16(r0);
                    if (c + 1 < c_end) issue(c + 1, r1, s1a, s1b);
                    process(c, r0, s0a, s0b);
                    if (c + 1 < c_end) {
                        tc_wait_ld16(r1);
                        if (c + 2 < c_end) issue(c + 2, r0, s0a, s0b);
                        process(c + 1, r1, s1a, s1b);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_remote(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
            }
            if (g.dbg) {
                dbg_wait += te1 - te0;
                dbg_proc += clock64() - te1;
            }
        }
        if (TMA_OUT && st_elected) tma_wait_group<0>();  // shared memory must outlive the last store
        if (g.dbg && warp == 2 && lane == 0) {
            g.dbg[blockIdx.x * 8 + 4] = dbg_wait;
            g.dbg[blockIdx.x * 8 + 5] = dbg_proc;
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------

int make_operand_tmap(CUtensorMap* tm, const GemmOperand& o, int box_rows_or_bw, int box_bh, int pix_stride = 1) {
    uint64_t dims[4];
    uint64_t strides[3];
    uint32_t box[4];
    uint32_t estr[4] = {1, 1, 1, 1};
    if (!o.conv) {
        dims[0] = static_cast<uint64_t>(o.inner);
        dims[1] = static_cast<uint64_t>(o.rows);
        dims[2] = static_cast<uint64_t>(o.nb2 > 0 ? o.nb2 : 1);
        dims[3] = static_cast<uint64_t>(o.nb1 > 0 ? o.nb1 : 1);
        strides[0] = static_cast<uint64_t>(o.row_stride) * 2;
        strides[1] = static_cast<uint64_t>(dims[2] > 1 ? o.b2_stride : o.row_stride * o.rows) * 2;
        strides[2] = static_cast<uint64_t>(dims[3] > 1 ? o.b1_stride : o.row_stride * o.rows) * 2;
        box[0] = 64;
        box[1] = static_cast<uint32_t>(box_rows_or_bw);
        box[2] = 1;
        box[3] = 1;
    } else {
        dims[0] = static_cast<uint64_t>(o.inner);
        dims[1] = static_cast<uint64_t>(o.W);
        dims[2] = static_cast<uint64_t>(o.H);
        dims[3] = static_cast<uint64_t>(o.nimg);
        strides[0] = static_cast<uint64_t>(o.row_stride) * 2;
        strides[1] = strides[0] * static_cast<uint64_t>(o.W);
        strides[2] = strides[1] * static_cast<uint64_t>(o.H);
        box[0] = 64;
        box[1] = static_cast<uint32_t>(box_rows_or_bw);
        box[2] = static_cast<uint32_t>(box_bh);
        box[3] = 1;
        if (pix_stride > 1) {
            // a box that lands bw x bh pixels taken every `pix_stride`-th pixel spans bw*stride x bh*stride tensor
            // elements: cuTensorMapEncodeTiled counts boxDim in traversed elements (verified on B200: the other reading
            // hangs on expect_tx)
            estr[1] = estr[2] = static_cast<uint32_t>(pix_stride);
            box[1] *= static_cast<uint32_t>(pix_stride);
            box[2] *= static_cast<uint32_t>(pix_stride);
        }
    }
    return encode_tmap(tm, o.ptr, 4, dims, strides, box, 0, estr);
}

// tile width (and CTA-group size) minimising waves x per-k16 cost.  `groups` = CTA groups that run concurrently.
int pick_bn(const GemmProblem& p, long long tiles_m_batches, int groups, int cg, int min_step = 1) {
    if (p.force_bn > 0) return p.force_bn;
    static int env_bn = -1;
    if (env_bn < 0) {
        const char* e_ = getenv("NK_GEMM_FORCE_BN");  // experiment switch
        env_bn = e_ ? atoi(e_) : 0;
    }
    if (env_bn > 0) return env_bn;
    int step = p.B.mn_major ? 64 * cg : 16 * cg;
    while (step % min_step != 0) step += p.B.mn_major ? 64 * cg : 16 * cg;  // TMA-store epilogue: whole column groups
    long long best_cost = -1;
    int best = 256;
    for (int bn = 256; bn >= step; bn -= step) {
        const long long tiles = tiles_m_batches * ((p.N + bn - 1) / bn);
        const long long waves = (tiles + groups - 1) / groups;
        // cycles per k16 step: BN/2 on the tensor pipe, or the L2->smem operand feed: per k-iteration every SM pulls a
        // 16 KB A slab plus BN/cg rows of B, and ~32 KB per 512 clk is what the fabric sustains when all SMs stream
        // (measured: profiles/r01_gemm_debug_timing.txt, tools/ktest mainloop_* cases), so narrow tiles are fed
        // no faster than wide ones do math and only pay off when they remove whole waves
        const long long per = std::max<long long>(bn / 2, 64 + bn / (2 * cg)) + 8;
        // with split-K (atomic accumulation) the K range is spread over idle CTAs, so total work counts, not waves
        const long long cost = (p.out == OUT_F32_ATOMIC) ? (tiles * per * 16) / groups : waves * per * 16;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

}  // namespace

// pixel-tile shape for a conv forward M tile (<=128 pixels) or wgrad K tile (exactly 64 pixels)
static void pick_pixel_tile(int H, int W, int target, bool exact, int* bw_out, int* bh_out) {
    long long best = -1;
    int bbw = 1, bbh = target;
    for (int bw = 1; bw <= std::min(W, target); ++bw) {
        if (exact && (target % bw) != 0) continue;
        int bh = target / bw;
        if (bh > 256) continue;
        if (bh > H) bh = exact ? bh : H;
        const long long tiles = static_cast<long long>((W + bw - 1) / bw) * ((H + bh - 1) / bh);
        if (best < 0 || tiles < best || (tiles == best && bw > bbw)) {
            best = tiles;
            bbw = bw;
            bbh = bh;
        }
    }
    *bw_out = bbw;
    *bh_out = bbh;
}

// debugging aid (NK_GEMM_DEBUG_TIMING): per-role cycle accounting of the launch that just ran
static int debug_report(const GemmProblem& p, const GemmDev& g, int cg, long long total, int nsm, long long* dbg_buf,
                        cudaStream_t stream) {
    NK_CUDA(cudaStreamSynchronize(stream));
    static long long host[8 * 1024];
    NK_CUDA(cudaMemcpy(host, dbg_buf, sizeof(host), cudaMemcpyDeviceToHost));
    const int nb = cg == 1 ? static_cast<int>(std::min<long long>(total, nsm))
                           : 2 * static_cast<int>(std::min<long long>(total, nsm / 2));
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int nlead = 0;
    for (int b = 0; b < nb && b < 1024; b += cg) {
        for (int i = 0; i < 8; ++i) a[i] += static_cast<double>(host[b * 8 + i]);
        ++nlead;
    }
    for (int i = 0; i < 8; ++i) a[i] /= std::max(1, nlead);
    fprintf(stderr,
            "[nk gemm dbg] mode=%d M=%d N=%d k_iters=%d BN=%d cg=%d stages=%d tiles=%lld splits=%d | per leader CTA: tiles %.1f  "
            "mma loop %.0f clk (wait full %.0f, wait tmem_empty %.0f)  producer wait empty %.0f  "
            "epilogue wait full %.0f proc %.0f\n",
            g.mode, p.M, p.N, g.k_iters, g.BN, cg, g.stages, total, g.splits, a[6], a[3], a[1], a[2], a[0], a[4], a[5]);
    return NK_OK;
}

// tensor map of an output matrix for the TMA-store epilogue: boxes of one 128-byte column group x the tile's rows
static int make_out_tmap(CUtensorMap* tm, const void* ptr, int is_f32, long long ncols, long long ld, const GemmDev& g,
                         const GemmProblem& p) {
    const uint64_t es = is_f32 ? 4 : 2;
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    dims[0] = static_cast<uint64_t>(ncols);
    strides[0] = static_cast<uint64_t>(ld) * es;
    box[0] = is_f32 ? 32 : 64;
    if (g.mode == MODE_CONV_FWD) {
        dims[1] = static_cast<uint64_t>(g.cW);
        dims[2] = static_cast<uint64_t>(g.cH);
        dims[3] = static_cast<uint64_t>(p.A.nimg);
        strides[1] = strides[0] * dims[1];
        strides[2] = strides[1] * dims[2];
        box[1] = static_cast<uint32_t>(g.bw);
        box[2] = static_cast<uint32_t>(g.bh);
        box[3] = 1;
    } else {
        dims[1] = static_cast<uint64_t>(p.M);
        dims[2] = static_cast<uint64_t>(g.nb2);
        dims[3] = static_cast<uint64_t>(g.nb1);
        strides[1] = (g.nb2 > 1 ? static_cast<uint64_t>(p.c_b2_stride) : static_cast<uint64_t>(ld) * dims[1]) * es;
        strides[2] = (g.nb1 > 1 ? static_cast<uint64_t>(p.c_b1_stride) : static_cast<uint64_t>(ld) * dims[1] * dims[2]) * es;
        box[1] = BM;
        box[2] = 1;
        box[3] = 1;
    }
    return encode_tmap(tm, ptr, 4, dims, strides, box, is_f32);
}

// Row-tile pairing (the DUAL instantiations): 0 = off, 1 = where the cost model below expects a gain, 2 = wherever legal.
// Starts from NK_GEMM_DUAL (default 0); neurosis_b200.tune switches it on at run time after the variant has reproduced
// the single-tile kernel's results bit for bit on the device it runs on (nk_gemm_set_dual).
static int g_dual_mode = -1;
static int dual_mode() {
    if (g_dual_mode < 0) {
        const char* e_ = getenv("NK_GEMM_DUAL");
        g_dual_mode = e_ ? atoi(e_) : 0;
    }
    return g_dual_mode;
}
int gemm_set_dual(int mode) {
    const int prev = dual_mode();
    if (mode >= 0 && mode <= 2) g_dual_mode = mode;
    return prev;
}
// mode 1 pairs only launches with at least this many 64-deep k-iterations: with a short reduction the two epilogues that
// pairing exposes per scheduler tile outweigh the saved operand traffic.  Where the break-even lies is measured by
// neurosis_b200.tune on the device (NK_GEMM_DUAL_MIN_K pins it; default 0 = no limit).
static int g_dual_min_k = -1;
static int dual_min_k() {
    if (g_dual_min_k < 0) {
        const char* e_ = getenv("NK_GEMM_DUAL_MIN_K");
        g_dual_min_k = e_ ? std::max(0, atoi(e_)) : 0;
    }
    return g_dual_min_k;
}
// DUAL: k-iterations by which the second row tile trails the first (see the MMA issuer); 0 = interleaved.  Clamped to
// stages - 1 by the kernel.  NK_GEMM_DUAL_SKEW pins it; default 0 until neurosis_b200.tune has measured it.
static int g_dual_skew = -1;
static int dual_skew() {
    if (g_dual_skew < 0) {
        const char* e_ = getenv("NK_GEMM_DUAL_SKEW");
        g_dual_skew = e_ ? std::max(0, std::min(7, atoi(e_))) : 0;
    }
    return g_dual_skew;
}
int gemm_set_dual_skew(int k_iters) {
    const int prev = dual_skew();
    if (k_iters >= 0 && k_iters <= 7) g_dual_skew = k_iters;
    return prev;
}
// L2 prefetch of the epilogue's side input at the start of every tile, a bit mask: bit 0 = h of EPI_GEGLU_BWD, bit 1 = the
// residual of EPI_LINEAR.  0 off (default, or NK_GEMM_EPI_PREFETCH).  A hint to the memory system, results are unchanged
// by construction.
static int g_epi_prefetch = -1;
static int epi_prefetch_mode() {
    if (g_epi_prefetch < 0) {
        const char* e_ = getenv("NK_GEMM_EPI_PREFETCH");
        g_epi_prefetch = e_ ? (atoi(e_) & 3) : 0;
    }
    return g_epi_prefetch;
}
int gemm_set_epi_prefetch(int on) {
    const int prev = epi_prefetch_mode();
    if (on >= 0 && on <= 3) g_epi_prefetch = on;
    return prev;
}
// Which launch classes may pair: bit 0 = matrix GEMM with K-major A (linear forward / data gradient), bit 1 = matrix GEMM
// with MN-major A (linear weight gradient), bit 2 = implicit-GEMM convolution forward / data gradient.  Default 7
// (NK_GEMM_DUAL_CLASSES); neurosis_b200.tune clears the bits of classes that lose on the device.
static int g_dual_classes = -1;
static int dual_classes() {
    if (g_dual_classes < 0) {
        const char* e_ = getenv("NK_GEMM_DUAL_CLASSES");
        g_dual_classes = e_ ? (atoi(e_) & 7) : 7;
    }
    return g_dual_classes;
}
int gemm_set_dual_classes(int mask) {
    const int prev = dual_classes();
    if (mask >= 0 && mask <= 7) g_dual_classes = mask;
    return prev;
}
int gemm_set_dual_min_k(int k_iters) {
    const int prev = dual_min_k();
    if (k_iters >= 0) g_dual_min_k = k_iters;
    return prev;
}

int launch_gemm(const GemmProblem& p, cudaStream_t stream) {
    static int nsm = 0;
    if (nsm == 0) nsm = device_sm_count();
    NK_REQUIRE(nsm > 0, NK_ERR_CUDA, "no CUDA device");
    NK_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, NK_ERR_SHAPE, "gemm: empty problem %d %d %d", p.M, p.N, p.K);

    GemmDev g;
    memset(&g, 0, sizeof(g));
    g.M = p.M;
    g.N = p.N;
    g.nb2 = p.nb2 > 0 ? p.nb2 : 1;
    g.nb1 = p.nb1 > 0 ? p.nb1 : 1;
    g.a_mn = p.A.mn_major;
    g.b_mn = p.B.mn_major;
    g.ksize = p.ksize > 0 ? p.ksize : 1;
    g.pad = p.pad;
    g.mode = MODE_PLAIN;
    if (p.wgrad) {
        g.mode = MODE_CONV_WGRAD;
        NK_REQUIRE(p.A.conv && p.B.conv && p.A.mn_major && p.B.mn_major, NK_ERR_UNSUPPORTED,
                   "wgrad needs MN-major image operands");
    } else if (p.A.conv) {
        g.mode = MODE_CONV_FWD;
        NK_REQUIRE(!p.A.mn_major && !p.B.conv && !p.B.mn_major, NK_ERR_UNSUPPORTED,
                   "conv forward needs K-major operands");
        NK_REQUIRE(p.A.inner % 64 == 0, NK_ERR_SHAPE, "conv: C_in %lld not a multiple of 64", p.A.inner);
    }

    static int env_no_swap = -1;
    if (env_no_swap < 0) env_no_swap = getenv("NK_NO_CONV_SWAP") ? 1 : 0;  // A/B switch
    const bool swap_ab = g.mode == MODE_CONV_FWD && p.N > 64 && p.N <= 128 && p.conv_stride <= 1 && p.out == OUT_BF16 &&
                         p.epi == EPI_LINEAR && !env_no_swap && p.force_bn == 0 && p.force_cta_group == 0;
    int a_box1 = BM, a_box2 = 1, b_box2 = 1;
    g.cstride = p.conv_stride > 1 ? p.conv_stride : 1;
    g.pad_t = g.cstride > 1 ? p.pad_t : p.pad;
    g.pad_l = g.cstride > 1 ? p.pad_l : p.pad;
    if (g.mode == MODE_CONV_FWD) {
        g.cH = g.cstride > 1 ? p.out_H : p.A.H;
        g.cW = g.cstride > 1 ? p.out_W : p.A.W;
        NK_REQUIRE(g.cH > 0 && g.cW > 0, NK_ERR_SHAPE, "conv: empty output image");
        pick_pixel_tile(g.cH, g.cW, BM, false, &g.bw, &g.bh);
        g.tiles_w = (g.cW + g.bw - 1) / g.bw;
        g.tiles_h = (g.cH + g.bh - 1) / g.bh;
        g.tiles_m128 = p.A.nimg * g.tiles_w * g.tiles_h;
        g.cin_blocks = static_cast<int>(p.A.inner / 64);
        g.k_iters_total = g.ksize * g.ksize * g.cin_blocks;
        a_box1 = g.bw;
        a_box2 = g.bh;
        g.a_bytes = static_cast<uint32_t>(g.bw * g.bh * 128);
    } else if (g.mode == MODE_CONV_WGRAD) {
        g.cH = p.A.H;
        g.cW = p.A.W;
        pick_pixel_tile(g.cH, g.cW, 64, true, &g.bw, &g.bh);
        g.tiles_w = (g.cW + g.bw - 1) / g.bw;
        g.tiles_h = (g.cH + g.bh - 1) / g.bh;
        g.tiles_m128 = (p.M + BM - 1) / BM;
        g.k_iters_total = p.A.nimg * g.tiles_w * g.tiles_h;
        a_box1 = g.bw;
        a_box2 = g.bh;
        b_box2 = g.bh;
        g.a_bytes = 2 * ATOM_BYTES;
        g.cin_blocks = static_cast<int>(p.B.inner / 64);
        NK_REQUIRE(p.B.inner % 64 == 0, NK_ERR_SHAPE, "wgrad: C_in %lld not a multiple of 64", p.B.inner);
        NK_REQUIRE(p.N == g.ksize * g.ksize * static_cast<int>(p.B.inner), NK_ERR_SHAPE, "wgrad: N must be taps*C_in");
    } else {
        g.tiles_m128 = (p.M + BM - 1) / BM;
        g.k_iters_total = (p.K + BK - 1) / BK;
        a_box1 = p.A.mn_major ? 64 : BM;
        g.a_bytes = A_STAGE_BYTES;
    }

    // CTA-group size: pairs (cta_group::2, 256-row tiles, B tile split across the pair) whenever the problem has
    // at least two 128-row tiles and a wide enough N; single CTAs otherwise.
    static long long* dbg_buf = nullptr;
    static int dbg_on = -1;
    if (dbg_on < 0) {
        dbg_on = getenv("NK_GEMM_DEBUG_TIMING") ? 1 : 0;
        if (dbg_on) NK_CUDA(cudaMalloc(&dbg_buf, 8 * 1024 * sizeof(long long)));
    }
    if (swap_ab) {
        // ---- transposed convolution forward: M = Cout (<= 128), N = 256 pixels, single CTAs ----
        g.mode = MODE_CONV_FWD_T;
        pick_pixel_tile(g.cH, g.cW, 256, false, &g.bw, &g.bh);
        g.tiles_w = (g.cW + g.bw - 1) / g.bw;
        g.tiles_h = (g.cH + g.bh - 1) / g.bh;
        g.tiles_m128 = 1;
        g.tiles_m = 1;
        g.BN = 256;
        g.tiles_n = p.A.nimg * g.tiles_w * g.tiles_h;
        g.a_bytes = A_STAGE_BYTES;
        g.b_bytes = static_cast<uint32_t>(g.bw * g.bh * 128);
        g.k_iters = g.k_iters_total;
        g.splits = 1;
        int e2 = make_operand_tmap(&g.tmA, p.B, BM, 1);  // packed weights [Cout, taps*Cin]: box {64 k, 128 rows}
        if (e2) return e2;
        e2 = make_operand_tmap(&g.tmB, p.A, g.bw, g.bh, 1);  // NHWC image: box {64 ch, bw, bh}
        if (e2) return e2;
        g.C = p.C;
        g.ldc = p.ldc;
        g.out = p.out;
        g.epi = p.epi;
        g.alpha = p.alpha;
        g.bias = p.bias;
        g.bias_img = p.bias_img;
        g.rows_per_img = p.A.nimg;  // transposed mode: number of images (tiles past the last image load nothing)
        g.residual = p.residual;
        g.ldr = p.ldr;
        g.idesc = make_idesc_bf16(BM, 256, 0, 0);
        const int stage_bytes_t = A_STAGE_BYTES + 256 * 128;
        const int max_smem_t = 227 * 1024;
        g.stages = std::max(2, std::min((max_smem_t - 1024 - 256) / stage_bytes_t, 8));
        const int smem_t = g.stages * stage_bytes_t + 1024 + (2 * g.stages + 5) * 8;
        static bool attr_t_set = false;
        if (!attr_t_set) {
            NK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1, VAR_TRANSPOSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_t));
            attr_t_set = true;
        }
        g.dbg = dbg_buf;
        if (dbg_buf) NK_CUDA(cudaMemsetAsync(dbg_buf, 0, 8 * 1024 * sizeof(long long), stream));
        const int grid_t = static_cast<int>(std::min<long long>(g.tiles_n, nsm));
        gemm_tc_kernel<1, VAR_TRANSPOSED><<<grid_t, NUM_THREADS, smem_t, stream>>>(g);
        NK_CUDA(cudaGetLastError());
        if (dbg_buf) return debug_report(p, g, 1, g.tiles_n, nsm, dbg_buf, stream);
        return NK_OK;
    }
    static int env_cg = -1;
    if (env_cg < 0) {
        const char* e_ = getenv("NK_GEMM_CTA_GROUP");  // debugging / A-B switch: 1 or 2
        env_cg = e_ ? atoi(e_) : 0;
    }
    int cg = 2;
    if (p.force_cta_group == 1 || p.force_cta_group == 2) cg = p.force_cta_group;
    else if (env_cg == 1) cg = 1;
    else if (g.tiles_m128 < 2 || p.N < 32 || (p.B.mn_major && p.N < 128)) cg = 1;
    if (p.force_bn > 0 && (p.force_bn % (p.B.mn_major ? 64 * cg : 16 * cg)) != 0) cg = 1;
    g.tiles_m = (g.tiles_m128 + cg - 1) / cg;

    const long long tiles_mb = static_cast<long long>(g.tiles_m) * g.nb2 * g.nb1;
    const bool geglu_fwd = p.epi == EPI_GEGLU_FWD, geglu_bwd = p.epi == EPI_GEGLU_BWD;
    // TMA-store epilogue (NK_GEMM_TMA_OUT=0: A/B switch back to per-thread global stores): needs 16-byte aligned rows
    static int env_tma_out = -1;
    if (env_tma_out < 0) {
        const char* e_ = getenv("NK_GEMM_TMA_OUT");
        env_tma_out = e_ ? atoi(e_) : 1;
    }
    const int out_es = p.out == OUT_BF16 ? 2 : 4;
    const int gw_out = p.out == OUT_BF16 ? 64 : 32;
    auto aligned16 = [](const void* ptr, long long ld_bytes) {
        return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0 && (ld_bytes & 15) == 0;
    };
    bool tma_out = env_tma_out != 0 && (p.epi == EPI_LINEAR || geglu_fwd || geglu_bwd) &&
                   (g.mode == MODE_PLAIN || g.mode == MODE_CONV_FWD || g.mode == MODE_CONV_WGRAD) &&
                   (p.C == nullptr ? geglu_fwd : aligned16(p.C, p.ldc * out_es)) &&
                   (g.nb2 <= 1 || (p.c_b2_stride * out_es) % 16 == 0) && (g.nb1 <= 1 || (p.c_b1_stride * out_es) % 16 == 0) &&
                   (!geglu_fwd || (aligned16(p.C2, p.ldc2 * 2) && p.N % 64 == 0)) && (!geglu_bwd || p.N % 64 == 0) &&
                   (p.force_bn <= 0 || p.force_bn % ((geglu_fwd ? 2 : 1) * gw_out) == 0);
    g.BN = pick_bn(p, tiles_mb, nsm / cg, cg, tma_out ? gw_out : 1);
    if (tma_out && g.BN % gw_out != 0) tma_out = false;  // NK_GEMM_FORCE_BN experiments
    if (geglu_fwd || geglu_bwd) {
        NK_REQUIRE(g.mode == MODE_PLAIN && !p.A.mn_major && p.nb1 <= 1 && p.nb2 <= 1 && p.out == OUT_BF16, NK_ERR_UNSUPPORTED,
                   "GEGLU epilogues: plain bf16 GEMM only");
        NK_REQUIRE(p.geglu_d == p.N && p.N % 16 == 0 && p.ldc % 8 == 0, NK_ERR_SHAPE, "GEGLU: D=%d must be a multiple of 16", p.N);
        g.geglu_d = p.geglu_d;
    }
    if (geglu_fwd) {  // an N tile = BN/2 value columns + their BN/2 gate columns (see the producer)
        NK_REQUIRE(!p.B.mn_major && p.C2 != nullptr && p.ldc2 % 8 == 0, NK_ERR_SHAPE, "GEGLU fwd: K-major W and an output");
        g.BN = p.N >= 128 ? 256 : 2 * p.N;
        g.C2 = p.C2;
        g.ldc2 = p.ldc2;
    }
    if (geglu_bwd) NK_REQUIRE(p.aux != nullptr && p.ldr % 8 == 0, NK_ERR_SHAPE, "GEGLU bwd needs h (aux)");
    NK_REQUIRE(g.BN >= 16 * cg && g.BN <= 256 && g.BN % (16 * cg) == 0, NK_ERR_SHAPE, "bad BN %d (cta group %d)", g.BN, cg);
    NK_REQUIRE(!p.B.mn_major || g.BN % (64 * cg) == 0, NK_ERR_SHAPE, "MN-major B needs BN %% %d == 0", 64 * cg);
    g.tiles_n = geglu_fwd ? (p.N + g.BN / 2 - 1) / (g.BN / 2) : (p.N + g.BN - 1) / g.BN;
    const int bnc = g.BN / cg;
    g.b_bytes = static_cast<uint32_t>(bnc) * 128u;

    // ---- row-tile pairing (DUAL) ----
    bool dual = false;
    long long tiles_mb_sched = tiles_mb;  // M tiles the scheduler walks (x batches)
    {
        const int dm = p.force_dual > 0 ? p.force_dual : (p.force_dual < 0 ? 0 : dual_mode());
        // (the batched attention GEMMs with their softmax epilogues stay unpaired: neurosis_b200.tune does not cover them)
        const bool legal = cg == 2 && p.epi == EPI_LINEAR && (g.mode == MODE_PLAIN || g.mode == MODE_CONV_FWD) &&
                           g.nb1 * g.nb2 == 1 && g.tiles_m >= 2;
        if (dm != 0 && legal) {
            const long long tm2 = (g.tiles_m + 1) / 2;
            const long long single = tiles_mb * g.tiles_n, paired = tm2 * g.nb2 * g.nb1 * g.tiles_n;
            const long long groups = nsm / cg;
            const int cls = g.mode == MODE_CONV_FWD ? 4 : (g.a_mn ? 2 : 1);
            if (dm == 2) {
                dual = true;
            } else if ((dual_classes() & cls) == 0) {
                dual = false;  // this class of launches lost on the device (neurosis_b200.tune)
            } else if (g.k_iters_total < dual_min_k()) {
                dual = false;  // reduction too short for pairing to pay (threshold measured by neurosis_b200.tune)
            } else if (p.out == OUT_F32_ATOMIC) {
                // split-K spreads the work over the machine either way: pairing pays unless the odd last tile wastes much
                dual = paired * 2 * 100 <= single * 115;
            } else {
                // feed-bound: a paired tile moves 48 KB per k-iteration where two single tiles move 64 KB (x 0.75)
                const long long ws = (single + groups - 1) / groups, wd = (paired + groups - 1) / groups;
                dual = wd * 2 * 75 < ws * 97;
            }
            if (dual) tiles_mb_sched = tm2 * g.nb2 * g.nb1;
            g.dual_skew = dual ? dual_skew() : 0;
        }
    }

    // split-K only when accumulating atomically: pick the split count that minimises
    // waves x (k-iterations per split + fixed per-tile cost), i.e. fills the last wave
    int splits = 1;
    if (p.out == OUT_F32_ATOMIC) {
        const long long tiles = tiles_mb_sched * g.tiles_n;
        const int groups = nsm / cg;
        static int env_splits = -1;
        if (env_splits < 0) {
            const char* e_ = getenv("NK_GEMM_FORCE_SPLITS");  // experiment switch
            env_splits = e_ ? atoi(e_) : 0;
        }
        if (p.force_splits > 0 || env_splits > 0) {
            splits = p.force_splits > 0 ? p.force_splits : env_splits;
        } else {
            const int max_s = std::max(1, std::min(64, g.k_iters_total / 4));
            long long best = -1;
            for (int sp = 1; sp <= max_s; ++sp) {
                const long long waves = (tiles * sp + groups - 1) / groups;
                const long long cost = waves * ((g.k_iters_total + sp - 1) / sp + 10);
                if (best < 0 || cost < best) {
                    best = cost;
                    splits = sp;
                }
            }
        }
        splits = std::max(1, std::min(splits, g.k_iters_total));
    }
    g.k_iters = (g.k_iters_total + splits - 1) / splits;
    g.splits = (g.k_iters_total + g.k_iters - 1) / g.k_iters;
    // Tile order.  The CTA groups that run at the same time should touch as few DISTINCT operand bytes as possible:
    // identical requests from different SMs are merged at the L2 (measured: with the A loads of 16384 x 10240 x 1280
    // removed — 40 distinct B tiles in flight — the kernel runs at 1229 TFLOP/s, with the B loads removed — 2 distinct A
    // slabs — at 1430, with neither at 1442; profiles/r02_gemm_probe_skip.txt), and an activation operand larger than
    // the 126 MB L2 must not be streamed from DRAM once per N tile (741 MB instead of 215 MB for 16384 x 1280 x 5120
    // in the first-round order).  So tiles are walked in groups of `gm` row blocks x all N tiles, row block fastest,
    // with gm ~ sqrt(#CTA groups): one wave covers ~gm row blocks x ~gm N tiles.  NK_GEMM_RASTER=0 restores the old
    // order (all row blocks of one N tile first), =1 is one row block at a time (A/B switches).
    {
        static int env_raster = -1;
        if (env_raster < 0) {
            const char* e_ = getenv("NK_GEMM_RASTER");
            env_raster = e_ ? atoi(e_) : 2;
        }
        if (dual) g.tiles_m = (g.tiles_m + 1) / 2;  // from here on: scheduler tiles (the kernel derives the row tiles)
        const int conc = nsm / cg;
        int side = 1;
        while (side * side < conc) ++side;
        const int ncols = std::max(1, std::min(g.tiles_n, side));
        int gm = (conc + ncols - 1) / ncols;
        if (env_raster == 0) gm = 0;
        else if (env_raster == 1) gm = 1;
        g.raster_gm = std::min(gm, g.tiles_m);
        static int env_skip = -1;
        if (env_skip < 0) {
            const char* e_ = getenv("NK_GEMM_DBG_SKIP");
            env_skip = e_ ? atoi(e_) : 0;
        }
        g.dbg_skip = (g.mode == MODE_PLAIN && !geglu_fwd) ? env_skip : 0;
    }

    g.a_b2 = (!p.A.conv && p.A.nb2 > 1) ? 1 : 0;
    g.a_b1 = (!p.A.conv && p.A.nb1 > 1) ? 1 : 0;
    g.b_b2 = (!p.B.conv && p.B.nb2 > 1) ? 1 : 0;
    g.b_b1 = (!p.B.conv && p.B.nb1 > 1) ? 1 : 0;

    int e = make_operand_tmap(&g.tmA, p.A, a_box1, a_box2, g.mode == MODE_CONV_FWD ? g.cstride : 1);
    if (e) return e;
    const int b_box1 = p.B.conv ? g.bw : (p.B.mn_major ? 64 : (geglu_fwd ? g.BN / 2 : bnc));
    e = make_operand_tmap(&g.tmB, p.B, b_box1, b_box2);
    if (e) return e;

    g.C = p.C;
    g.ldc = p.ldc;
    g.c_b2 = p.c_b2_stride;
    g.c_b1 = p.c_b1_stride;
    g.out = p.out;
    g.epi = p.epi;
    g.alpha = p.alpha;
    g.bias = p.bias;
    g.bias_img = p.bias_img;
    g.rows_per_img = p.rows_per_img > 0 ? p.rows_per_img : 1;
    g.residual = p.residual;
    g.ldr = p.ldr;
    g.rowvec = p.rowvec;
    g.aux = p.aux;
    g.idesc = make_idesc_bf16(BM * cg, g.BN, g.a_mn, g.b_mn);
    NK_REQUIRE(p.epi != EPI_DSOFTMAX || (p.aux && p.rowvec), NK_ERR_SHAPE, "dsoftmax needs aux+rowvec");
    NK_REQUIRE(p.epi != EPI_EXP2 || p.rowvec, NK_ERR_SHAPE, "exp2 epilogue needs rowvec");

    g.tma_out = tma_out ? 1 : 0;
    g.stg_per_half = geglu_fwd ? 3 : 2;
    g.has_c = p.C != nullptr;
    if (tma_out) {
        if (p.C != nullptr) {
            e = make_out_tmap(&g.tmC, p.C, p.out != OUT_BF16, (geglu_fwd || geglu_bwd) ? 2LL * p.N : p.N, p.ldc, g, p);
            if (e) return e;
        }
        if (geglu_fwd) {
            e = make_out_tmap(&g.tmC2, p.C2, 0, p.N, p.ldc2, g, p);
            if (e) return e;
        } else if (epi_prefetch_mode()) {
            // side-input map (same logical shape and batch strides as the output, its own row stride)
            if (geglu_bwd && (epi_prefetch_mode() & 1) && aligned16(p.aux, p.ldr * 2)) {
                e = make_out_tmap(&g.tmC2, p.aux, 0, 2LL * p.N, p.ldr, g, p);
                if (e) return e;
                g.epi_prefetch = 1;
            } else if (!geglu_bwd && (epi_prefetch_mode() & 2) && p.epi == EPI_LINEAR && p.residual != nullptr &&
                       aligned16(p.residual, p.ldr * 2)) {
                e = make_out_tmap(&g.tmC2, p.residual, 0, p.N, p.ldr, g, p);
                if (e) return e;
                g.epi_prefetch = 1;
            }
        }
    }
    const int staging_bytes = tma_out ? 2 * g.stg_per_half * STG_BYTES : 0;
    const int stage_bytes = (dual ? 2 : 1) * A_STAGE_BYTES + bnc * 128;
    const int max_smem = 227 * 1024;
    int stages = (max_smem - 1024 - 256 - staging_bytes) / stage_bytes;
    stages = std::max(2, std::min(stages, 8));
    g.stages = stages;
    const int smem_bytes = stages * stage_bytes + staging_bytes + 1024 /*align slack*/ + (2 * stages + 5) * 8;
    const long long total = tiles_mb_sched * g.tiles_n * g.splits;
    NK_REQUIRE(total < (1LL << 31), NK_ERR_SHAPE, "too many tiles");
    g.dbg = dbg_buf;
    if (dbg_buf) NK_CUDA(cudaMemsetAsync(dbg_buf, 0, 8 * 1024 * sizeof(long long), stream));
    // kernel instantiation: CTA-group size x epilogue family x TMA-store epilogue
    const void* fn = nullptr;
    {
        const int var = geglu_fwd ? VAR_GEGLU_FWD : (geglu_bwd ? VAR_GEGLU_BWD : VAR_MAIN);
#define NK_PICK(CGV, VARV, TMAV) \
    if (cg == CGV && var == VARV && tma_out == TMAV && !dual) fn = reinterpret_cast<const void*>(&gemm_tc_kernel<CGV, VARV, TMAV>);
        if (dual && !tma_out) fn = reinterpret_cast<const void*>(&gemm_tc_kernel<2, VAR_MAIN, false, true>);
        if (dual && tma_out) fn = reinterpret_cast<const void*>(&gemm_tc_kernel<2, VAR_MAIN, true, true>);
        NK_PICK(1, VAR_MAIN, false) NK_PICK(2, VAR_MAIN, false) NK_PICK(1, VAR_MAIN, true) NK_PICK(2, VAR_MAIN, true)
        NK_PICK(1, VAR_GEGLU_FWD, false) NK_PICK(2, VAR_GEGLU_FWD, false) NK_PICK(1, VAR_GEGLU_FWD, true)
        NK_PICK(2, VAR_GEGLU_FWD, true) NK_PICK(1, VAR_GEGLU_BWD, false) NK_PICK(2, VAR_GEGLU_BWD, false)
        NK_PICK(1, VAR_GEGLU_BWD, true) NK_PICK(2, VAR_GEGLU_BWD, true)
#undef NK_PICK
    }
    NK_REQUIRE(fn != nullptr, NK_ERR_UNSUPPORTED, "no kernel instantiation");
    {
        static const void* attr_done[16];
        static int n_attr_done = 0;
        bool seen = false;
        for (int i = 0; i < n_attr_done; ++i) seen = seen || attr_done[i] == fn;
        if (!seen) {
            NK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
            if (n_attr_done < 16) attr_done[n_attr_done++] = fn;
        }
    }
    {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        const int grid = cg == 1 ? static_cast<int>(std::min<long long>(total, nsm))
                                 : 2 * static_cast<int>(std::min<long long>(total, nsm / 2));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cg;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        void* args[1] = {const_cast<GemmDev*>(&g)};
        NK_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    }
    NK_CUDA(cudaGetLastError());
    if (dbg_buf) return debug_report(p, g, cg, total, nsm, dbg_buf, stream);
    return NK_OK;
}

}  // namespace nk
