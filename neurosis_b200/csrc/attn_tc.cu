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

namespace nk {
namespace {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int HD = 64;
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int ATT_THREADS = 192;

struct alignas(64) AttnDev {
    CUtensorMap tmQ, tmK, tmV;
    bf16* O;
    float* lse;
    long long o_row_stride, o_batch_stride;  // elements; head offset = h*HD
    int Nq, Nk, H, B;
    float scale_log2;  // softmax scale * log2(e)
    float scale;
};

// smem layout (after 1024B alignment): Q | K[3] | V[3] | P[2][2] | barriers
constexpr int SM_Q = 0;
constexpr int SM_K = SM_Q + TILE_BYTES;
constexpr int SM_V = SM_K + KV_STAGES * TILE_BYTES;
constexpr int SM_P = SM_V + KV_STAGES * TILE_BYTES;
constexpr int SM_BAR = SM_P + 4 * TILE_BYTES;
constexpr int ATT_SMEM = SM_BAR + 256 + 1024;

__global__ void __launch_bounds__(ATT_THREADS, 1) attn_fwd_kernel(const __grid_constant__ AttnDev g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* q_full = bars;                 // 1
    uint64_t* kv_full = bars + 1;            // KV_STAGES
    uint64_t* kv_empty = bars + 1 + KV_STAGES;
    uint64_t* s_full = bars + 1 + 2 * KV_STAGES;  // 2
    uint64_t* p_full = s_full + 2;                // 2 (128 arrivals)
    uint64_t* o_full = p_full + 2;                // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

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
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 128);
            mbar_init(&o_full[i], 1);
        }
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
    const uint32_t TM_S = 0, TM_O = 256;  // S buffers at cols 0,128 ; O partial buffers at 256,320

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
            auto issue_s = [&](int j, int stage) {
                const uint64_t k_desc = make_smem_desc(smem_u32(smem + SM_K + stage * TILE_BYTES), 16u, 1024u);
                const uint32_t d = tmem_base + TM_S + static_cast<uint32_t>((j & 1) * 128);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    tc_mma_ss(d, q_desc + static_cast<uint64_t>(k * 2), k_desc + static_cast<uint64_t>(k * 2), idesc_s,
                              k > 0 ? 1u : 0u);
                tc_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0, 20);
            int stage = 0, nstage = 0;
            uint32_t nphase = 0;
            mbar_wait(&kv_full[0], 0, 21);
            tc_fence_after();
            issue_s(0, 0);
            nstage = 1 % KV_STAGES;
            nphase = (KV_STAGES == 1) ? 1u : 0u;
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) {
                    mbar_wait(&kv_full[nstage], nphase, 22u);
                    tc_fence_after();
                    issue_s(j + 1, nstage);
                }
                mbar_wait(&p_full[j & 1], (j >> 1) & 1u, 23u);
                tc_fence_after();
                {
                    const uint32_t d = tmem_base + TM_O + static_cast<uint32_t>((j & 1) * 64);
                    const uint32_t pbase = smem_u32(smem + SM_P + (j & 1) * 2 * TILE_BYTES);
                    const uint32_t vbase = smem_u32(smem + SM_V + stage * TILE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BKV / 16; ++kk) {
                        const uint64_t p_desc =
                            make_smem_desc(pbase + static_cast<uint32_t>((kk >> 2) * TILE_BYTES + (kk & 3) * 32), 16u, 1024u);
                        const uint64_t v_desc = make_smem_desc(vbase + static_cast<uint32_t>(kk * 2048), 8192u, 1024u);
                        tc_mma_ss(d, p_desc, v_desc, idesc_o, kk > 0 ? 1u : 0u);
                    }
                    tc_commit(&o_full[j & 1]);
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
        for (int j = 0; j < T; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1u, 30u);
            tc_fence_after();
            const uint32_t s_addr = lane_addr + TM_S + static_cast<uint32_t>((j & 1) * 128);
            const int kv_valid = g.Nk - j * BKV;  // columns >= kv_valid are padding
            // pass 1: row max
            float m_tile = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t raw[32];
                tc_ld32(s_addr + static_cast<uint32_t>(c * 32), raw);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = (c * 32 + i < kv_valid) ? __uint_as_float(raw[i]) : -INFINITY;
                    m_tile = fmaxf(m_tile, v);
                }
            }
            const float m_new = fmaxf(m_run, m_tile);
            const float alpha = exp2f((m_run - m_new) * g.scale_log2);  // 0 on the first tile
            const float mb = m_new * g.scale_log2;
            // pass 2: probabilities -> smem (bf16, swizzled), row sum
            float l_tile = 0.f;
            uint8_t* pbuf = smem + SM_P + (j & 1) * 2 * TILE_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t raw[32];
                tc_ld32(s_addr + static_cast<uint32_t>(c * 32), raw);
                tc_wait_ld();
                float p[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = exp2f(__uint_as_float(raw[i]) * g.scale_log2 - mb);
                    p[i] = (c * 32 + i < kv_valid) ? v : 0.f;
                    l_tile += p[i];
                }
                // columns c*32 .. c*32+31 -> K-block (c>>1), 16B chunks ((c&1)*4 .. +3)
                uint8_t* rowp = pbuf + (c >> 1) * TILE_BYTES + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 w;
                    w.x = pack_bf16x2(p[8 * q + 0], p[8 * q + 1]);
                    w.y = pack_bf16x2(p[8 * q + 2], p[8 * q + 3]);
                    w.z = pack_bf16x2(p[8 * q + 4], p[8 * q + 5]);
                    w.w = pack_bf16x2(p[8 * q + 6], p[8 * q + 7]);
                    const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
                    *reinterpret_cast<uint4*>(rowp + chunk * 16) = w;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&p_full[j & 1]);
            // fold in the previous tile's partial product, then rescale to the new max
            if (j > 0) {
                mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1u, 31u);
                tc_fence_after();
                const uint32_t o_addr = lane_addr + TM_O + static_cast<uint32_t>(((j - 1) & 1) * 64);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t raw[32];
                    tc_ld32(o_addr + static_cast<uint32_t>(c * 32), raw);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[c * 32 + i] = (o[c * 32 + i] + __uint_as_float(raw[i])) * alpha;
                }
            }
            l_run = l_run * alpha + l_tile;
            m_run = m_new;
        }
        // last partial product
        mbar_wait(&o_full[(T - 1) & 1], ((T - 1) >> 1) & 1u, 32u);
        tc_fence_after();
        {
            const uint32_t o_addr = lane_addr + TM_O + static_cast<uint32_t>(((T - 1) & 1) * 64);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tc_ld32(o_addr + static_cast<uint32_t>(c * 32), raw);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(raw[i]);
            }
        }
        const int q = q0 + r;
        if (q < g.Nq) {
            const float inv = 1.f / l_run;
            bf16* op = g.O + static_cast<long long>(b) * g.o_batch_stride + static_cast<long long>(q) * g.o_row_stride +
                       static_cast<long long>(h) * HD;
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
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
        tmem_dealloc(tmem_base, 512);
    }
}

int make_qkv_tmap(CUtensorMap* tm, const void* p, int N, int H, int B, long long row_stride, long long head_stride,
                  long long batch_stride) {
    const uint64_t dims[4] = {static_cast<uint64_t>(HD), static_cast<uint64_t>(H), static_cast<uint64_t>(N),
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
    NK_REQUIRE(head_dim == HD, NK_ERR_UNSUPPORTED, "fused attention supports head_dim 64 (got %d)", head_dim);
    NK_REQUIRE(B > 0 && H > 0 && Nq > 0 && Nk > 0, NK_ERR_SHAPE, "attention: empty problem");
    NK_REQUIRE(o_row_stride % 8 == 0 && o_batch_stride % 8 == 0, NK_ERR_SHAPE, "attention: output strides");
    AttnDev g;
    memset(&g, 0, sizeof(g));
    int e = make_qkv_tmap(&g.tmQ, q, Nq, H, B, q_row_stride, HD, q_batch_stride);
    if (e) return e;
    e = make_qkv_tmap(&g.tmK, k, Nk, H, B, k_row_stride, HD, k_batch_stride);
    if (e) return e;
    e = make_qkv_tmap(&g.tmV, v, Nk, H, B, v_row_stride, HD, v_batch_stride);
    if (e) return e;
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

}  // extern "C"
