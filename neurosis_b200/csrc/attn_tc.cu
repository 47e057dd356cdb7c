// Fused (flash-style) attention forward on tcgen05 / TMEM, head_dim 64, no mask, no dropout.
//
//   O = softmax(scale * Q K^T) V      per (batch, head), tensors laid out [B, N, H, D]
//
// One CTA handles 128 query rows of one (batch, head):
//   warp 0     TMA producer: Q tile once, then a ring of (K_j, V_j) 128-row tiles
//   warp 1     MMA issuer:   S_j = Q K_j^T (TMEM, double buffered)   and   O_j = P_j V_j (TMEM, x2)
//   warps 2-5  softmax:      one thread per query row; reads S_j from TMEM, online max/sum,
//                            writes P_j (bf16, 128B-swizzled K-major) to smem for the PV MMA,
//                            accumulates O in registers from the per-tile partial products.
// S_{j+1} is issued before P_j V_j so the tensor pipe works while the softmax warps run.
//
// Replaces F.scaled_dot_product_attention / xformers.memory_efficient_attention at
// /root/reference/src/neurosis/modules/attention.py:346-352,410-412 (self and cross attention).
#include "common.cuh"

#include <cstdlib>

namespace nk {
namespace {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int HD = 64;
constexpr int KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int ATT_THREADS = 192;
constexpr int ATT_BWD_THREADS = 320;  // TMA warp, MMA warp, 8 softmax warps (2 per TMEM lane quadrant)

struct alignas(64) AttnDev {
    CUtensorMap tmQ, tmK, tmV;
    bf16* O;
    float* lse;
    long long o_row_stride, o_batch_stride;  // elements; head offset = h*HD
    int Nq, Nk, H, B;
    int D;             // valid head dim (<= 64, multiple of 8): columns D..63 of the Q/K/V tiles are zero-filled by TMA
    int p_tmem;        // 1: P stays in tensor memory (A operand of the PV MMA); 0: swizzled shared-memory tile (NK_ATTN_P_TMEM=0)
    float scale_log2;  // softmax scale * log2(e)
    float scale;
};

// smem layout (after 1024B alignment): Q | K[2] | V[2] | P[2 atoms] | barriers  (= 113 KB: two CTAs per SM, whose
// serial S -> softmax -> PV chains interleave on the tensor and MUFU pipes)
constexpr int SM_Q = 0;
constexpr int SM_K = SM_Q + TILE_BYTES;
constexpr int SM_V = SM_K + KV_STAGES * TILE_BYTES;
constexpr int SM_P = SM_V + KV_STAGES * TILE_BYTES;
constexpr int SM_BAR = SM_P + 2 * TILE_BYTES;
// two CTAs per SM need 2 * (ATT_SMEM + 1 KB reserved) <= 228 KB, which leaves 896 bytes of alignment slack; the
// dynamic window starts 1024-aligned in practice (no static smem) and the kernel traps if it ever does not fit
constexpr int ATT_SMEM = 115712;
static_assert(SM_BAR + 128 <= ATT_SMEM, "attention forward smem budget");
constexpr int ATT_TMEM_COLS = 256;  // S at 0..127, O partial at 128..191

// ---- softmax building blocks shared by the forward kernel: one thread = one query row, S row lives in TMEM ----
template <bool MASK>
__device__ __forceinline__ float tile_rowmax(uint32_t s_addr, int kv_valid) {
    float m0 = -INFINITY, m1 = -INFINITY;
    uint32_t ra[16], rb[16];
    tc_ld16(s_addr, ra);
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
        tc_wait_ld16(ra);
        tc_ld16(s_addr + static_cast<uint32_t>((c + 1) * 16), rb);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float a = __uint_as_float(ra[i]);
            if (MASK) a = (c * 16 + i < kv_valid) ? a : -INFINITY;
            if (i & 1) m1 = fmaxf(m1, a); else m0 = fmaxf(m0, a);
        }
        tc_wait_ld16(rb);
        if (c + 2 < 8) tc_ld16(s_addr + static_cast<uint32_t>((c + 2) * 16), ra);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float a = __uint_as_float(rb[i]);
            if (MASK) a = ((c + 1) * 16 + i < kv_valid) ? a : -INFINITY;
            if (i & 1) m1 = fmaxf(m1, a); else m0 = fmaxf(m0, a);
        }
    }
    return fmaxf(m0, m1);
}

// p = exp2(s*scale_log2 - mb) -> bf16, written to the 128B-swizzled [row][kv] tile pair at pbuf; returns the row sum
// p_tmem != 0: the bf16 probabilities go to tensor memory (columns p_tmem + 8c .. + 8: two per 32-bit column) as the
// A operand of the PV MMA, instead of the swizzled shared-memory tile at pbuf
template <bool MASK>
__device__ __forceinline__ float tile_probs(uint32_t s_addr, int kv_valid, float scale_log2, float mb, uint8_t* pbuf,
                                            int r, uint32_t p_tmem = 0) {
    float l0 = 0.f, l1 = 0.f;
    auto emit = [&](int c, const uint32_t (&raw)[16]) {  // c = 16-column chunk 0..7
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            float v0 = exp2f(fmaf(__uint_as_float(raw[i]), scale_log2, -mb));
            float v1 = exp2f(fmaf(__uint_as_float(raw[i + 1]), scale_log2, -mb));
            if (MASK) {
                v0 = (c * 16 + i < kv_valid) ? v0 : 0.f;
                v1 = (c * 16 + i + 1 < kv_valid) ? v1 : 0.f;
            }
            l0 += v0;
            l1 += v1;
            w[i >> 1] = pack_bf16x2(v0, v1);
        }
        if (p_tmem != 0) {
            tc_st8(p_tmem + static_cast<uint32_t>(c * 8), w);
            return;
        }
        uint8_t* rowp = pbuf + (c >> 2) * TILE_BYTES + r * 128;
        const int ch = (c & 3) * 2;
        *reinterpret_cast<uint4*>(rowp + ((ch ^ (r & 7)) * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) * 16)) = make_uint4(w[4], w[5], w[6], w[7]);
    };
    uint32_t ra[16], rb[16];
    tc_ld16(s_addr, ra);
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
        tc_wait_ld16(ra);
        tc_ld16(s_addr + static_cast<uint32_t>((c + 1) * 16), rb);
        emit(c, ra);
        tc_wait_ld16(rb);
        if (c + 2 < 8) tc_ld16(s_addr + static_cast<uint32_t>((c + 2) * 16), ra);
        emit(c + 1, rb);
    }
    return l0 + l1;
}

__global__ void __launch_bounds__(ATT_THREADS, 2) attn_fwd_kernel(const __grid_constant__ AttnDev g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    if (static_cast<int>(smem - smem_raw) + SM_BAR + 128 > ATT_SMEM) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* q_full = bars;                 // 1
    uint64_t* kv_full = bars + 1;            // KV_STAGES
    uint64_t* kv_empty = bars + 1 + KV_STAGES;
    uint64_t* s_full = bars + 1 + 2 * KV_STAGES;  // S_j complete in TMEM
    uint64_t* p_full = s_full + 1;                // P_j in smem and S_j consumed (128 arrivals)
    uint64_t* o_full = p_full + 1;                // P_j V_j complete in TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const int T = (g.Nk + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&g.tmQ);
        tma_prefetch_desc(&g.tmK);
        tma_prefetch_desc(&g.tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_full, 128);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, ATT_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t TM_S = 0, TM_O = 128, TM_P = 192;  // P: 128 x 128 bf16 = 64 columns

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tma_load_4d(&g.tmQ, q_full, smem + SM_Q, 0, h, q0, b);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < T; ++j) {
                mbar_wait(&kv_empty[stage], phase ^ 1u, 10u + stage);
                mbar_arrive_expect_tx(&kv_full[stage], 2 * TILE_BYTES);
                tma_load_4d(&g.tmK, &kv_full[stage], smem + SM_K + stage * TILE_BYTES, 0, h, j * BKV, b);
                tma_load_4d(&g.tmV, &kv_full[stage], smem + SM_V + stage * TILE_BYTES, 0, h, j * BKV, b);
                if (++stage == KV_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);  // S: both K-major
            const uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);   // O: P K-major, V MN-major
            const uint64_t q_desc = make_smem_desc(smem_u32(smem + SM_Q), 16u, 1024u);
            auto issue_s = [&](int stage) {
                const uint64_t k_desc = make_smem_desc(smem_u32(smem + SM_K + stage * TILE_BYTES), 16u, 1024u);
                const uint32_t d = tmem_base + TM_S;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    tc_mma_ss(d, q_desc + static_cast<uint64_t>(k * 2), k_desc + static_cast<uint64_t>(k * 2), idesc_s,
                              k > 0 ? 1u : 0u);
                tc_commit(s_full);
            };
            mbar_wait(q_full, 0, 20);
            int stage = 0, nstage = 0;
            uint32_t nphase = 0;
            mbar_wait(&kv_full[0], 0, 21);
            tc_fence_after();
            issue_s(0);
            nstage = 1 % KV_STAGES;
            nphase = (KV_STAGES == 1) ? 1u : 0u;
            for (int j = 0; j < T; ++j) {
                // P_j written and S_j fully read by the softmax warps (they also folded O_{j-1} before writing P_j)
                mbar_wait(p_full, j & 1u, 23u);
                tc_fence_after();
                if (j + 1 < T) {
                    mbar_wait(&kv_full[nstage], nphase, 22u);
                    tc_fence_after();
                    issue_s(nstage);
                }
                {
                    const uint32_t d = tmem_base + TM_O;
                    const uint32_t pbase = smem_u32(smem + SM_P);
                    const uint32_t vbase = smem_u32(smem + SM_V + stage * TILE_BYTES);
                    if (g.p_tmem) {  // A = P from tensor memory: 8 columns (16 bf16) per k16 step
#pragma unroll
                        for (int kk = 0; kk < BKV / 16; ++kk) {
                            const uint64_t v_desc = make_smem_desc(vbase + static_cast<uint32_t>(kk * 2048), 8192u, 1024u);
                            tc_mma_ts(d, tmem_base + TM_P + static_cast<uint32_t>(kk * 8), v_desc, idesc_o, kk > 0 ? 1u : 0u);
                        }
                    } else {
#pragma unroll
                        for (int kk = 0; kk < BKV / 16; ++kk) {
                            const uint64_t p_desc = make_smem_desc(
                                pbase + static_cast<uint32_t>((kk >> 2) * TILE_BYTES + (kk & 3) * 32), 16u, 1024u);
                            const uint64_t v_desc = make_smem_desc(vbase + static_cast<uint32_t>(kk * 2048), 8192u, 1024u);
                            tc_mma_ss(d, p_desc, v_desc, idesc_o, kk > 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(o_full);
                    tc_commit(&kv_empty[stage]);
                }
                stage = nstage;
                if (++nstage == KV_STAGES) {
                    nstage = 0;
                    nphase ^= 1u;
                }
            }
        }
    } else {
        // ===================== softmax / output warps =====================
        const int quad = warp & 3;
        const int r = quad * 32 + lane;  // query row within the tile == TMEM lane
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        float o[HD];
#pragma unroll
        for (int i = 0; i < HD; ++i) o[i] = 0.f;
        float m_run = -INFINITY;  // running max of raw scores
        float l_run = 0.f;
        // o / l_run lag one tile behind: at tile j they are expressed relative to max m_{j-2}; folding in tile j-1's
        // partial product (computed relative to m_{j-1}) is then ONE fma per element with alpha_{j-1} = 2^(m_{j-2}-m_{j-1})
        float alpha_prev = 0.f, l_prev = 0.f;
        for (int j = 0; j < T; ++j) {
            mbar_wait(s_full, j & 1u, 30u);
            tc_fence_after();
            const uint32_t s_addr = lane_addr + TM_S;
            const int kv_valid = g.Nk - j * BKV;  // columns >= kv_valid are padding
            // pass 1: row max ; pass 2: probabilities -> smem (bf16, swizzled) + row sum.  Only the last K/V tile
            // can be partial, so the masked variant is taken at most once per row.
            const bool tail = kv_valid < BKV;
            const float m_tile = tail ? tile_rowmax<true>(s_addr, kv_valid) : tile_rowmax<false>(s_addr, kv_valid);
            const float m_new = fmaxf(m_run, m_tile);
            const float alpha = exp2f((m_run - m_new) * g.scale_log2);  // 0 on the first tile
            const float mb = m_new * g.scale_log2;
            // fold in the previous tile's partial product; its completion also means the tensor pipe is done reading
            // the P buffer, which is rewritten next
            if (j > 0) {
                mbar_wait(o_full, (j - 1) & 1u, 31u);
                tc_fence_after();
                const uint32_t o_addr = lane_addr + TM_O;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t raw[32];
                    tc_ld32(o_addr + static_cast<uint32_t>(c * 32), raw);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(raw[i]));
                }
                l_run = fmaf(l_run, alpha_prev, l_prev);
            }
            uint8_t* pbuf = smem + SM_P;
            const uint32_t p_dst = g.p_tmem ? lane_addr + TM_P : 0u;
            l_prev = tail ? tile_probs<true>(s_addr, kv_valid, g.scale_log2, mb, pbuf, r, p_dst)
                          : tile_probs<false>(s_addr, kv_valid, g.scale_log2, mb, pbuf, r, p_dst);
            if (g.p_tmem) tc_wait_st(); else fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
            alpha_prev = alpha;
            m_run = m_new;
        }
        // last partial product
        mbar_wait(o_full, (T - 1) & 1u, 32u);
        tc_fence_after();
        {
            const uint32_t o_addr = lane_addr + TM_O;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tc_ld32(o_addr + static_cast<uint32_t>(c * 32), raw);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(raw[i]));
            }
            l_run = fmaf(l_run, alpha_prev, l_prev);
        }
        const int q = q0 + r;
        if (q < g.Nq) {
            const float inv = 1.f / l_run;
            bf16* op = g.O + static_cast<long long>(b) * g.o_batch_stride + static_cast<long long>(q) * g.o_row_stride +
                       static_cast<long long>(h) * g.D;
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
                if (8 * c >= g.D) break;  // head dims below 64 (SD1.5: 40): only the valid columns exist in memory
                uint4 w;
                w.x = pack_bf16x2(o[8 * c + 0] * inv, o[8 * c + 1] * inv);
                w.y = pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
                w.z = pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
                w.w = pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
                reinterpret_cast<uint4*>(op)[c] = w;
            }
            if (g.lse) g.lse[(static_cast<long long>(b) * g.H + h) * g.Nq + q] = m_run * g.scale + logf(l_run);
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}


// =================================================================================================
// Wide-head forward (head_dim = 64 * NC, NC <= 8; the VAE mid-block attention is one head of 512).
// S = Q K^T is accumulated in TMEM over the NC 64-wide chunks of the head dimension; the CTA owns ONE 64-wide chunk
// of V / O (blockIdx.y = head * NC + v-chunk), so the online-softmax state is the same 64 registers per row as in
// the head_dim-64 kernel.  The NC CTAs of a head recompute the same S: redundant tensor work (idle otherwise) instead
// of the [Nq, Nk] fp32 score matrix in HBM that the materialised path needs.
//   smem: Q resident (NC x 16 KB) | K chunk ring (2 x 16 KB) | V ring (2 x 16 KB) | P (32 KB)      (224 KB at NC = 8)
//   TMEM: S double-buffered (2 x 128 columns) | O partial (64 columns)
// =================================================================================================
struct alignas(64) AttnWideDev {
    CUtensorMap tmQ, tmK, tmV;  // dims {64, H*NC, N, B}: "chunk-heads"
    bf16* O;
    float* lse;
    long long o_row_stride, o_batch_stride;
    int Nq, Nk, H, B, NC;
    float scale_log2, scale;
};

__global__ void __launch_bounds__(ATT_THREADS, 1) attn_fwd_wide_kernel(const __grid_constant__ AttnWideDev g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    const int NC = g.NC;
    uint8_t* sm_q = smem;
    uint8_t* sm_k = sm_q + NC * TILE_BYTES;
    uint8_t* sm_v = sm_k + 2 * TILE_BYTES;
    uint8_t* sm_p = sm_v + 2 * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + 2 * TILE_BYTES);
    uint64_t* q_full = bars;        // 1
    uint64_t* k_full = bars + 1;    // 2
    uint64_t* k_empty = bars + 3;   // 2
    uint64_t* v_full = bars + 5;    // 2
    uint64_t* v_empty = bars + 7;   // 2
    uint64_t* s_full = bars + 9;    // 2
    uint64_t* p_full = bars + 11;   // 1 (128 arrivals)
    uint64_t* o_full = bars + 12;   // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ;
    const int h = blockIdx.y / NC;
    const int vc = blockIdx.y - h * NC;  // the V / O chunk of this CTA
    const int b = blockIdx.z;
    const int T = (g.Nk + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&g.tmQ);
        tma_prefetch_desc(&g.tmK);
        tma_prefetch_desc(&g.tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1);
        }
        mbar_init(p_full, 128);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t TM_S = 0, TM_O = 256;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, static_cast<uint32_t>(NC) * TILE_BYTES);
            for (int c = 0; c < NC; ++c) tma_load_4d(&g.tmQ, q_full, sm_q + c * TILE_BYTES, 0, h * NC + c, q0, b);
            int ks = 0;
            uint32_t kph = 0;
            for (int j = 0; j < T; ++j) {
                for (int c = 0; c < NC; ++c) {
                    mbar_wait(&k_empty[ks], kph ^ 1u, 10u + ks);
                    mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
                    tma_load_4d(&g.tmK, &k_full[ks], sm_k + ks * TILE_BYTES, 0, h * NC + c, j * BKV, b);
                    if (++ks == 2) {
                        ks = 0;
                        kph ^= 1u;
                    }
                }
                const int vs = j & 1;
                mbar_wait(&v_empty[vs], ((j >> 1) & 1u) ^ 1u, 12u + vs);
                mbar_arrive_expect_tx(&v_full[vs], TILE_BYTES);
                tma_load_4d(&g.tmV, &v_full[vs], sm_v + vs * TILE_BYTES, 0, h * NC + vc, j * BKV, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
            const uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);
            int ks = 0;
            uint32_t kph = 0;
            auto issue_s = [&](int j) {
                const uint32_t d = tmem_base + TM_S + static_cast<uint32_t>((j & 1) * 128);
                for (int c = 0; c < NC; ++c) {
                    mbar_wait(&k_full[ks], kph, 22u);
                    tc_fence_after();
                    const uint64_t q_desc = make_smem_desc(smem_u32(sm_q + c * TILE_BYTES), 16u, 1024u);
                    const uint64_t k_desc = make_smem_desc(smem_u32(sm_k + ks * TILE_BYTES), 16u, 1024u);
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        tc_mma_ss(d, q_desc + static_cast<uint64_t>(k * 2), k_desc + static_cast<uint64_t>(k * 2), idesc_s,
                                  (c > 0 || k > 0) ? 1u : 0u);
                    tc_commit(&k_empty[ks]);
                    if (++ks == 2) {
                        ks = 0;
                        kph ^= 1u;
                    }
                }
                tc_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0, 20);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < T; ++j) {
                // S_{j+1} goes to the other TMEM buffer (its last reader, the softmax of tile j-1, finished before
                // p_full(j-1), which this warp has already waited for) and overlaps the softmax of tile j
                if (j + 1 < T) issue_s(j + 1);
                mbar_wait(p_full, j & 1u, 23u);
                const int vs = j & 1;
                mbar_wait(&v_full[vs], (j >> 1) & 1u, 24u);
                tc_fence_after();
                const uint32_t d = tmem_base + TM_O;
                const uint32_t pbase = smem_u32(sm_p);
                const uint32_t vbase = smem_u32(sm_v + vs * TILE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BKV / 16; ++kk) {
                    const uint64_t p_desc =
                        make_smem_desc(pbase + static_cast<uint32_t>((kk >> 2) * TILE_BYTES + (kk & 3) * 32), 16u, 1024u);
                    const uint64_t v_desc = make_smem_desc(vbase + static_cast<uint32_t>(kk * 2048), 8192u, 1024u);
                    tc_mma_ss(d, p_desc, v_desc, idesc_o, kk > 0 ? 1u : 0u);
                }
                tc_commit(o_full);
                tc_commit(&v_empty[vs]);
            }
        }
    } else {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        float o[HD];
#pragma unroll
        for (int i = 0; i < HD; ++i) o[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f, l_prev = 0.f;
        for (int j = 0; j < T; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1u, 30u);
            tc_fence_after();
            const uint32_t s_addr = lane_addr + TM_S + static_cast<uint32_t>((j & 1) * 128);
            const int kv_valid = g.Nk - j * BKV;
            const bool tail = kv_valid < BKV;
            const float m_tile = tail ? tile_rowmax<true>(s_addr, kv_valid) : tile_rowmax<false>(s_addr, kv_valid);
            const float m_new = fmaxf(m_run, m_tile);
            const float alpha = exp2f((m_run - m_new) * g.scale_log2);
            const float mb = m_new * g.scale_log2;
            if (j > 0) {  // also: P_{j-1} V_{j-1} is done reading the single P buffer
                mbar_wait(o_full, (j - 1) & 1u, 31u);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t raw[32];
                    tc_ld32(lane_addr + TM_O + static_cast<uint32_t>(c * 32), raw);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(raw[i]));
                }
                l_run = fmaf(l_run, alpha_prev, l_prev);
            }
            l_prev = tail ? tile_probs<true>(s_addr, kv_valid, g.scale_log2, mb, sm_p, r)
                          : tile_probs<false>(s_addr, kv_valid, g.scale_log2, mb, sm_p, r);
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
            alpha_prev = alpha;
            m_run = m_new;
        }
        mbar_wait(o_full, (T - 1) & 1u, 32u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t raw[32];
            tc_ld32(lane_addr + TM_O + static_cast<uint32_t>(c * 32), raw);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(raw[i]));
        }
        l_run = fmaf(l_run, alpha_prev, l_prev);
        const int q = q0 + r;
        if (q < g.Nq) {
            const float inv = 1.f / l_run;
            bf16* op = g.O + static_cast<long long>(b) * g.o_batch_stride + static_cast<long long>(q) * g.o_row_stride +
                       static_cast<long long>(h * NC + vc) * HD;
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
                uint4 w;
                w.x = pack_bf16x2(o[8 * c + 0] * inv, o[8 * c + 1] * inv);
                w.y = pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
                w.z = pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
                w.w = pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
                reinterpret_cast<uint4*>(op)[c] = w;
            }
            if (g.lse && vc == 0) g.lse[(static_cast<long long>(b) * g.H + h) * g.Nq + q] = m_run * g.scale + logf(l_run);
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =================================================================================================
// Fused attention backward (head_dim 64).  One CTA owns one 128-row K/V tile of one (batch, head) and
// loops over the 128-row Q tiles:
//     S  = Q_i K^T            dP = dO_i V^T                      (TMEM, 128x128 each)
//     P  = exp2(scale*log2e*S - lse*log2e)                        (softmax warps, bf16 -> smem)
//     dS = P * (dP - delta) * scale                               (softmax warps, bf16 -> smem)
//     dV += P^T dO_i          dK += dS^T Q_i                      (TMEM accumulators over the Q loop)
//     dQ_i = dS K             -> red.global.add.f32 into dQ_acc   (other K/V tiles add to the same rows)
// P and dS are stored once in the 128B-swizzled [q][kv] layout and consumed both as K-major (dQ) and as
// MN-major (dV, dK) UMMA operands; dO, Q and K are likewise consumed in both majors without transposes.
// =================================================================================================
struct alignas(64) AttnBwdDev {
    CUtensorMap tmQ, tmK, tmV, tmdO;
    CUtensorMap tmdQ;    // fp32 [B, Nq, H, 64] accumulator, box {32 cols, 1, 128 rows, 1}: target of the TMA reduce-add
    const float* lse;    // [B, H, Nq]
    const float* delta;  // [B, H, Nq]
    float* dQ;           // fp32 accumulator [B, Nq, H, 64] (zero-initialised by the caller)
    bf16* dK;
    bf16* dV;            // [B, Nk, H, 64]
    long long dq_row_stride, dq_batch_stride, dkv_row_stride, dkv_batch_stride;
    int Nq, Nk, H, B;
    int D;    // valid head dim (<= 64, multiple of 8); the tiles' columns D..63 are TMA zero fill
    float scale, scale_log2;
    int dbg;  // NK_ATTN_DBG A/B switch: 4 = per-thread red.global for dQ instead of the staged TMA reduce-add
};

constexpr int BW_K = 0;
constexpr int BW_V = BW_K + TILE_BYTES;
constexpr int BW_Q = BW_V + TILE_BYTES;        // 2 stages
constexpr int BW_DO = BW_Q + 2 * TILE_BYTES;   // 2 stages
constexpr int BW_P = BW_DO + 2 * TILE_BYTES;   // 128x128 bf16 = 2 tiles
constexpr int BW_DS = BW_P + 2 * TILE_BYTES;
constexpr int BW_DQ = BW_DS + 2 * TILE_BYTES;  // fp32 dQ staging: 2 x (128 rows x 32 cols), 128B-swizzled like the rest
constexpr int BW_BAR = BW_DQ + 2 * TILE_BYTES;
constexpr int BW_SMEM = BW_BAR + 256 + 1024;

__global__ void __launch_bounds__(ATT_BWD_THREADS, 1) attn_bwd_kernel(const __grid_constant__ AttnBwdDev g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW_BAR);
    uint64_t* kv_full = bars;            // 1
    uint64_t* qdo_full = bars + 1;       // 2
    uint64_t* qdo_empty = bars + 3;      // 2
    uint64_t* s_full = bars + 5;
    uint64_t* dp_full = bars + 6;
    uint64_t* p_ready = bars + 7;        // 256 arrivals
    uint64_t* ds_ready = bars + 8;       // 256 arrivals
    uint64_t* dq_full = bars + 9;
    uint64_t* dv_done = bars + 10;       // commit after the dV MMAs: P smem may be rewritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kv0 = blockIdx.x * BKV;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const int Tq = (g.Nq + BQ - 1) / BQ;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&g.tmQ);
        tma_prefetch_desc(&g.tmK);
        tma_prefetch_desc(&g.tmV);
        tma_prefetch_desc(&g.tmdO);
        tma_prefetch_desc(&g.tmdQ);
        mbar_init(kv_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&qdo_full[s], 1);
            mbar_init(&qdo_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(dp_full, 1);
        mbar_init(p_ready, 256);
        mbar_init(ds_ready, 256);
        mbar_init(dq_full, 1);
        mbar_init(dv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
            tma_load_4d(&g.tmK, kv_full, smem + BW_K, 0, h, kv0, b);
            tma_load_4d(&g.tmV, kv_full, smem + BW_V, 0, h, kv0, b);
            for (int i = 0; i < Tq; ++i) {
                const int st = i & 1;
                mbar_wait(&qdo_empty[st], (((i >> 1) & 1) ^ 1) & 1u, 40u + st);
                mbar_arrive_expect_tx(&qdo_full[st], 2 * TILE_BYTES);
                tma_load_4d(&g.tmQ, &qdo_full[st], smem + BW_Q + st * TILE_BYTES, 0, h, i * BQ, b);
                tma_load_4d(&g.tmdO, &qdo_full[st], smem + BW_DO + st * TILE_BYTES, 0, h, i * BQ, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_kk = make_idesc_bf16(128, 128, 0, 0);  // S, dP: both operands K-major
            const uint32_t idesc_mm = make_idesc_bf16(128, HD, 1, 1);   // dV, dK: both operands MN-major
            const uint32_t idesc_km = make_idesc_bf16(128, HD, 0, 1);   // dQ: dS K-major, K MN-major
            const uint32_t k_addr = smem_u32(smem + BW_K), v_addr = smem_u32(smem + BW_V);
            const uint32_t p_addr = smem_u32(smem + BW_P), ds_addr = smem_u32(smem + BW_DS);
            auto issue_s_dp = [&](int i) {
                const int st = i & 1;
                const uint32_t q_addr = smem_u32(smem + BW_Q + st * TILE_BYTES);
                const uint32_t do_addr = smem_u32(smem + BW_DO + st * TILE_BYTES);
                mbar_wait(&qdo_full[st], (i >> 1) & 1u, 51u);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k)  // S = Q K^T   (K = d = 64: 4 k-steps of 32 bytes)
                    tc_mma_ss(tmem_base + TM_S, make_smem_desc(q_addr + k * 32, 16u, 1024u),
                              make_smem_desc(k_addr + k * 32, 16u, 1024u), idesc_kk, k > 0 ? 1u : 0u);
                tc_commit(s_full);
#pragma unroll
                for (int k = 0; k < 4; ++k)  // dP = dO V^T
                    tc_mma_ss(tmem_base + TM_DP, make_smem_desc(do_addr + k * 32, 16u, 1024u),
                              make_smem_desc(v_addr + k * 32, 16u, 1024u), idesc_kk, k > 0 ? 1u : 0u);
                tc_commit(dp_full);
            };
            mbar_wait(kv_full, 0, 50);
            issue_s_dp(0);
            for (int i = 0; i < Tq; ++i) {
                const int st = i & 1;
                const uint32_t par = static_cast<uint32_t>(i & 1);
                const uint32_t q_addr = smem_u32(smem + BW_Q + st * TILE_BYTES);
                const uint32_t do_addr = smem_u32(smem + BW_DO + st * TILE_BYTES);
                // dV += P^T dO   (M = kv, N = d, K = q = 128: 8 k-steps of 16 q-rows = 2048 bytes)
                mbar_wait(p_ready, par, 52u);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_ss(tmem_base + TM_DV, make_smem_desc(p_addr + k * 2048, 16384u, 1024u),
                              make_smem_desc(do_addr + k * 2048, 8192u, 1024u), idesc_mm, (i > 0 || k > 0) ? 1u : 0u);
                tc_commit(dv_done);
                // dK += dS^T Q ;  dQ_i = dS K
                mbar_wait(ds_ready, par, 53u);
                tc_fence_after();
                // S/dP of the next tile go ahead of dK/dQ (their TMEM is free: the softmax warps passed ds_ready(i)), so
                // the softmax warps find S_{i+1} ready when they come back from the dQ flush
                if (i + 1 < Tq) issue_s_dp(i + 1);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_ss(tmem_base + TM_DK, make_smem_desc(ds_addr + k * 2048, 16384u, 1024u),
                              make_smem_desc(q_addr + k * 2048, 8192u, 1024u), idesc_mm, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_ss(tmem_base + TM_DQ,
                              make_smem_desc(ds_addr + (k >> 2) * TILE_BYTES + (k & 3) * 32, 16u, 1024u),
                              make_smem_desc(k_addr + k * 2048, 8192u, 1024u), idesc_km, k > 0 ? 1u : 0u);
                tc_commit(dq_full);
                tc_commit(&qdo_empty[st]);
            }
        }
    } else {
        // ===================== softmax warps: 2 per TMEM lane quadrant, each owning 64 of the 128 kv columns =====
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;  // kv columns [64*half, 64*half + 64) == 128B-swizzled smem tile `half`
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const int kv_valid = g.Nk - kv0;  // columns >= kv_valid are padding
        const uint32_t p_row = smem_u32(smem + BW_P) + static_cast<uint32_t>(half * TILE_BYTES + r * 128);
        const uint32_t ds_row = smem_u32(smem + BW_DS) + static_cast<uint32_t>(half * TILE_BYTES + r * 128);
        const bool kv_full_tile = kv_valid >= BKV;
        const int col0 = half * 64;
        // dQ partial of tile `it` (this warp: 32 of the 64 columns) -> global fp32 accumulator
        // The 128x64 fp32 partial is staged in smem (one 128x32 swizzled tile per column half) and added to global
        // memory by ONE TMA reduce per half: per-thread red.global (32 scattered half-sectors per warp instruction)
        // cost a quarter of the kernel.  Rows past Nq are clipped by the TMA unit.
        const uint32_t dq_row = smem_u32(smem + BW_DQ) + static_cast<uint32_t>(half * TILE_BYTES + r * 128);
        const bool dq_issuer = (warp == 2 && lane == 0);
        auto flush_dq = [&](int it) {
            mbar_wait(dq_full, static_cast<uint32_t>(it & 1), 62u);
            tc_fence_after();
            uint32_t raw[32];
            tc_ld32(lane_addr + TM_DQ + static_cast<uint32_t>(half * 32), raw);
            tc_wait_ld();
            tc_fence_before();
            if (g.dbg & 4) {
                const int q = it * BQ + r;
                float* dqp = g.dQ + static_cast<long long>(b) * g.dq_batch_stride +
                             static_cast<long long>(q) * g.dq_row_stride + static_cast<long long>(h) * HD + half * 32;
                if (q < g.Nq) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dqp + 4 * k),
                                     "f"(__uint_as_float(raw[4 * k])), "f"(__uint_as_float(raw[4 * k + 1])),
                                     "f"(__uint_as_float(raw[4 * k + 2])), "f"(__uint_as_float(raw[4 * k + 3]))
                                     : "memory");
                }
                return;
            }
            if (dq_issuer) tma_wait_group_read<0>();  // the previous tile's reduce has finished reading the staging tile
            named_bar_sync(1, 256);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                st_shared_v4(dq_row + static_cast<uint32_t>((k ^ (r & 7)) * 16),
                             make_uint4(raw[4 * k], raw[4 * k + 1], raw[4 * k + 2], raw[4 * k + 3]));
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (dq_issuer) {
                tma_reduce_add_4d(&g.tmdQ, smem + BW_DQ, 0, h, it * BQ, b);
                tma_reduce_add_4d(&g.tmdQ, smem + BW_DQ + TILE_BYTES, 32, h, it * BQ, b);
                tma_commit_group();
            }
        };
        const long long sbase = (static_cast<long long>(b) * g.H + h) * g.Nq;
        float lse_next = (r < g.Nq) ? g.lse[sbase + r] : 0.f;
        float dlt_next = (r < g.Nq) ? g.delta[sbase + r] : 0.f;
        for (int i = 0; i < Tq; ++i) {
            const uint32_t par = static_cast<uint32_t>(i & 1);
            const int q = i * BQ + r;
            const bool valid = q < g.Nq;
            const float lse2 = lse_next * 1.4426950408889634f;
            const float dlt_s = dlt_next * g.scale;  // dS = P * (dP - delta) * scale = P * fma(dP, scale, -delta*scale)
            {  // prefetch the next tile's row statistics: their L2 latency hides behind this tile
                const int qn = q + BQ;
                const bool vn = (i + 1 < Tq) && (qn < g.Nq);
                lse_next = vn ? g.lse[sbase + qn] : 0.f;
                dlt_next = vn ? g.delta[sbase + qn] : 0.f;
            }
            const bool nomask = kv_full_tile && ((i + 1) * BQ <= g.Nq);  // warp-uniform: whole tile in range
            // ---- P = exp2(scale*log2e*S - lse*log2e) ----
            mbar_wait(s_full, par, 60u);
            tc_fence_after();
            uint32_t ra[32], rb[32];
            tc_ld32(lane_addr + TM_S + static_cast<uint32_t>(col0), ra);
            tc_ld32(lane_addr + TM_S + static_cast<uint32_t>(col0 + 32), rb);
            if (i > 0) mbar_wait(dv_done, static_cast<uint32_t>((i - 1) & 1), 63u);  // dV_{i-1} no longer reads P
            tc_wait_ld();
            auto emit_p = [&](int cc, const float (&p)[32]) {
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    uint4 w;
                    w.x = pack_bf16x2(p[8 * qq + 0], p[8 * qq + 1]);
                    w.y = pack_bf16x2(p[8 * qq + 2], p[8 * qq + 3]);
                    w.z = pack_bf16x2(p[8 * qq + 4], p[8 * qq + 5]);
                    w.w = pack_bf16x2(p[8 * qq + 6], p[8 * qq + 7]);
                    st_shared_v4(p_row + static_cast<uint32_t>(((cc * 4 + qq) ^ (r & 7)) * 16), w);
                }
            };
            if (nomask) {  // warp-uniform: no per-element predicate in the common case
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    float p[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        p[k] = exp2f(fmaf(__uint_as_float(cc ? rb[k] : ra[k]), g.scale_log2, -lse2));
                    emit_p(cc, p);
                }
            } else {
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    float p[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const float e = exp2f(fmaf(__uint_as_float(cc ? rb[k] : ra[k]), g.scale_log2, -lse2));
                        p[k] = (valid && (col0 + cc * 32 + k < kv_valid)) ? e : 0.f;
                    }
                    emit_p(cc, p);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_ready);
            // ---- dQ of the previous tile (its MMAs ran behind our P pass); also frees the dS smem tile ----
            if (i > 0) flush_dq(i - 1);
            // ---- dS = P * (dP - delta) * scale ; P (bf16) is read back from this thread's own smem row ----
            mbar_wait(dp_full, par, 61u);
            tc_fence_after();
            tc_ld32(lane_addr + TM_DP + static_cast<uint32_t>(col0), ra);
            tc_ld32(lane_addr + TM_DP + static_cast<uint32_t>(col0 + 32), rb);
            uint4 pw[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) pw[j] = ld_shared_v4(p_row + static_cast<uint32_t>((j ^ (r & 7)) * 16));
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // 16-byte chunk j = columns 8j .. 8j+7 of this warp's 64
                const uint32_t pin[4] = {pw[j].x, pw[j].y, pw[j].z, pw[j].w};
                uint32_t dout[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = 8 * j + 2 * e;  // 0..63
                    const float dp0 = __uint_as_float(col < 32 ? ra[col & 31] : rb[col & 31]);
                    const float dp1 = __uint_as_float(col + 1 < 32 ? ra[(col + 1) & 31] : rb[(col + 1) & 31]);
                    // t = (dP - delta) * scale in fp32, rounded once to bf16; dS = P * t as ONE packed bf16x2 multiply
                    // (P is already bf16 and is 0 where masked): 4 issue slots per pair instead of 7
                    const uint32_t t2 = pack_bf16x2(fmaf(dp0, g.scale, -dlt_s), fmaf(dp1, g.scale, -dlt_s));
                    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(dout[e]) : "r"(pin[e]), "r"(t2));
                }
                st_shared_v4(ds_row + static_cast<uint32_t>((j ^ (r & 7)) * 16), make_uint4(dout[0], dout[1], dout[2], dout[3]));
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(ds_ready);
        }
        flush_dq(Tq - 1);
        if (dq_issuer) tma_wait_group<0>();  // smem must outlive the last reduce
        // ---- dV, dK of this K/V tile (all MMAs retired: dq_full of the last tile covers them); 32 columns per warp ----
        const int kvrow = kv0 + r;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            uint32_t raw[32];
            tc_ld32(lane_addr + (which ? TM_DK : TM_DV) + static_cast<uint32_t>(half * 32), raw);
            tc_wait_ld();
            if (kvrow < g.Nk) {
                bf16* op = (which ? g.dK : g.dV) + static_cast<long long>(b) * g.dkv_batch_stride +
                           static_cast<long long>(kvrow) * g.dkv_row_stride + static_cast<long long>(h) * g.D + half * 32;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (half * 32 + 8 * c >= g.D) break;
                    uint4 w;
                    w.x = pack_bf16x2(__uint_as_float(raw[8 * c + 0]), __uint_as_float(raw[8 * c + 1]));
                    w.y = pack_bf16x2(__uint_as_float(raw[8 * c + 2]), __uint_as_float(raw[8 * c + 3]));
                    w.z = pack_bf16x2(__uint_as_float(raw[8 * c + 4]), __uint_as_float(raw[8 * c + 5]));
                    w.w = pack_bf16x2(__uint_as_float(raw[8 * c + 6]), __uint_as_float(raw[8 * c + 7]));
                    reinterpret_cast<uint4*>(op)[c] = w;
                }
            }
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int make_qkv_tmap(CUtensorMap* tm, const void* p, int N, int H, int B, long long row_stride, long long head_stride,
                  long long batch_stride, int d_valid = HD) {
    // d_valid < 64: the 64-wide box reaches past the head; the TMA unit fills the missing columns with zeros
    const uint64_t dims[4] = {static_cast<uint64_t>(d_valid), static_cast<uint64_t>(H), static_cast<uint64_t>(N),
                              static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(head_stride) * 2, static_cast<uint64_t>(row_stride) * 2,
                                 static_cast<uint64_t>(batch_stride) * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    return encode_tmap_bf16(tm, p, 4, dims, strides, box);
}

}  // namespace
}  // namespace nk

using namespace nk;

extern "C" {

// q: [B, Nq, H, 64] with element strides (q_batch_stride, q_row_stride, 64 per head); k, v likewise with Nk.
// o: [B, Nq, H, 64] (o_row_stride between tokens); lse: fp32 [B, H, Nq] or null.
int nk_attention_fwd(const void* q, int64_t q_row_stride, int64_t q_batch_stride, const void* k, int64_t k_row_stride,
                     int64_t k_batch_stride, const void* v, int64_t v_row_stride, int64_t v_batch_stride, void* o,
                     int64_t o_row_stride, int64_t o_batch_stride, float* lse, int B, int H, int Nq, int Nk,
                     int head_dim, float scale, nk_stream_t stream) {
    ::nk::enter(stream);
    NK_REQUIRE((head_dim % HD == 0 && head_dim >= HD && head_dim <= 8 * HD) || (head_dim < HD && head_dim % 8 == 0 && head_dim >= 8),
               NK_ERR_UNSUPPORTED, "fused attention supports head_dim 8..64 in steps of 8 and 64..512 in steps of 64 (got %d)",
               head_dim);
    NK_REQUIRE(B > 0 && H > 0 && Nq > 0 && Nk > 0, NK_ERR_SHAPE, "attention: empty problem");
    NK_REQUIRE(o_row_stride % 8 == 0 && o_batch_stride % 8 == 0, NK_ERR_SHAPE, "attention: output strides");
    if (head_dim > HD) {  // wide heads: heads must be packed (head stride = head_dim) so that 64-wide chunks tile them
        const int NC = head_dim / HD;
        AttnWideDev w;
        memset(&w, 0, sizeof(w));
        int e = make_qkv_tmap(&w.tmQ, q, Nq, H * NC, B, q_row_stride, HD, q_batch_stride);
        if (e) return e;
        e = make_qkv_tmap(&w.tmK, k, Nk, H * NC, B, k_row_stride, HD, k_batch_stride);
        if (e) return e;
        e = make_qkv_tmap(&w.tmV, v, Nk, H * NC, B, v_row_stride, HD, v_batch_stride);
        if (e) return e;
        w.O = static_cast<bf16*>(o);
        w.lse = lse;
        w.o_row_stride = o_row_stride;
        w.o_batch_stride = o_batch_stride;
        w.Nq = Nq;
        w.Nk = Nk;
        w.H = H;
        w.B = B;
        w.NC = NC;
        w.scale = scale;
        w.scale_log2 = scale * 1.4426950408889634f;
        const int smem = (NC + 6) * TILE_BYTES + 256 + 1024;
        static bool wide_attr_set = false;
        if (!wide_attr_set) {
            NK_CUDA(cudaFuncSetAttribute(attn_fwd_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (8 + 6) * TILE_BYTES + 256 + 1024));
            wide_attr_set = true;
        }
        dim3 grid((Nq + BQ - 1) / BQ, H * NC, B);
        attn_fwd_wide_kernel<<<grid, ATT_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(w);
        NK_CUDA(cudaGetLastError());
        return NK_OK;
    }
    AttnDev g;
    memset(&g, 0, sizeof(g));
    int e = make_qkv_tmap(&g.tmQ, q, Nq, H, B, q_row_stride, head_dim, q_batch_stride, head_dim);
    if (e) return e;
    e = make_qkv_tmap(&g.tmK, k, Nk, H, B, k_row_stride, head_dim, k_batch_stride, head_dim);
    if (e) return e;
    e = make_qkv_tmap(&g.tmV, v, Nk, H, B, v_row_stride, head_dim, v_batch_stride, head_dim);
    if (e) return e;
    g.D = head_dim;
    {
        static int env_p = -1;
        if (env_p < 0) {
            const char* e_ = getenv("NK_ATTN_P_TMEM");
            env_p = e_ ? atoi(e_) : 1;
        }
        g.p_tmem = env_p;
    }
    g.O = static_cast<bf16*>(o);
    g.lse = lse;
    g.o_row_stride = o_row_stride;
    g.o_batch_stride = o_batch_stride;
    g.Nq = Nq;
    g.Nk = Nk;
    g.H = H;
    g.B = B;
    g.scale = scale;
    g.scale_log2 = scale * 1.4426950408889634f;
    static bool attr_set = false;
    if (!attr_set) {
        NK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set = true;
    }
    dim3 grid((Nq + BQ - 1) / BQ, H, B);
    attn_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM, static_cast<cudaStream_t>(stream)>>>(g);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}


// Gradients of nk_attention_fwd.  dq_acc: fp32 [B, Nq, H, 64] ZERO-INITIALISED accumulator (contiguous);
// dk, dv: bf16 [B, Nk, H, 64] with row / batch strides dkv_*_stride (0 = contiguous), e.g. column slices of one
// fused [B, Nk, 3*H*64] gradient buffer.  lse from the forward, delta[b,h,q] = sum_d dO*O (nk_attn_delta).
int nk_attention_bwd(const void* q, int64_t q_row_stride, int64_t q_batch_stride, const void* k, int64_t k_row_stride,
                     int64_t k_batch_stride, const void* v, int64_t v_row_stride, int64_t v_batch_stride,
                     const void* dO, int64_t do_row_stride, int64_t do_batch_stride, const float* lse,
                     const float* delta, float* dq_acc, void* dk, void* dv, int64_t dkv_row_stride,
                     int64_t dkv_batch_stride, int B, int H, int Nq, int Nk, int head_dim, float scale,
                     nk_stream_t stream) {
    ::nk::enter(stream);
    NK_REQUIRE(head_dim <= HD && head_dim >= 8 && head_dim % 8 == 0, NK_ERR_UNSUPPORTED,
               "fused attention backward supports head_dim 8..64 in steps of 8 (got %d)", head_dim);
    NK_REQUIRE(B > 0 && H > 0 && Nq > 0 && Nk > 0, NK_ERR_SHAPE, "attention bwd: empty problem");
    AttnBwdDev g;
    memset(&g, 0, sizeof(g));
    const int D = head_dim;
    int e = make_qkv_tmap(&g.tmQ, q, Nq, H, B, q_row_stride, D, q_batch_stride, D);
    if (e) return e;
    e = make_qkv_tmap(&g.tmK, k, Nk, H, B, k_row_stride, D, k_batch_stride, D);
    if (e) return e;
    e = make_qkv_tmap(&g.tmV, v, Nk, H, B, v_row_stride, D, v_batch_stride, D);
    if (e) return e;
    e = make_qkv_tmap(&g.tmdO, dO, Nq, H, B, do_row_stride, D, do_batch_stride, D);
    if (e) return e;
    g.D = D;
    {
        const uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(H), static_cast<uint64_t>(Nq),
                                  static_cast<uint64_t>(B)};
        const uint64_t strides[3] = {static_cast<uint64_t>(D) * 4, static_cast<uint64_t>(H) * D * 4,
                                     static_cast<uint64_t>(Nq) * H * D * 4};
        const uint32_t box[4] = {32, 1, 128, 1};
        e = encode_tmap(&g.tmdQ, dq_acc, 4, dims, strides, box, 1);
        if (e) return e;
    }
    g.lse = lse;
    g.delta = delta;
    g.dQ = dq_acc;
    g.dK = static_cast<bf16*>(dk);
    g.dV = static_cast<bf16*>(dv);
    g.dq_row_stride = static_cast<long long>(H) * D;
    g.dq_batch_stride = static_cast<long long>(Nq) * H * D;
    g.dkv_row_stride = dkv_row_stride > 0 ? dkv_row_stride : static_cast<long long>(H) * D;
    g.dkv_batch_stride = dkv_batch_stride > 0 ? dkv_batch_stride : static_cast<long long>(Nk) * H * D;
    NK_REQUIRE(g.dkv_row_stride % 8 == 0 && g.dkv_batch_stride % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(dk) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dv) & 15u) == 0,
               NK_ERR_SHAPE, "attention bwd: dk/dv must be 16-byte aligned with strides that are multiples of 8");
    g.Nq = Nq;
    g.Nk = Nk;
    g.H = H;
    g.B = B;
    g.scale = scale;
    g.scale_log2 = scale * 1.4426950408889634f;
    static int dbg = -1;
    if (dbg < 0) {
        const char* e_ = getenv("NK_ATTN_DBG");
        dbg = e_ ? atoi(e_) : 0;
    }
    g.dbg = dbg;
    static bool attr_set = false;
    if (!attr_set) {
        NK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM));
        attr_set = true;
    }
    dim3 grid((Nk + BKV - 1) / BKV, H, B);
    attn_bwd_kernel<<<grid, ATT_BWD_THREADS, BW_SMEM, static_cast<cudaStream_t>(stream)>>>(g);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}

}  // extern "C"
