// Multi-tensor optimizer-side kernels: the Adafactor step of the reference's
// training configs and the EMA shadow update, each as a fixed handful of launches over ALL parameters instead of
// ~15 ATen kernels per parameter (1 680 parameters for SDXL).
//
// Reference:
//   optimizers/adafactor.py:13-256   Adafactor.step (factored second moments over the last two dims, relative step,
//                                    parameter-scale learning rate, update clipping, optional first moment)
//   modules/ema.py:40-59             LitEma.forward: shadow -= (1 - decay) * (shadow - param)
//
// HBM-bound fp32 streaming.  Algorithmic bytes per parameter element of one Adafactor step: pass 1 reads p and g (8 B),
// pass 2 reads g (4 B), pass 3 reads g and p and writes p (12 B) plus the bf16 mirror (2 B) = 26 B; the factored
// moments are O(rows + cols).  EMA: 12 B per element.
//
// A parameter is described once (device table, built by the host) as [Bt, R, C] = (prod(shape[:-2]), shape[-2],
// shape[-1]):
//   KIND_VEC   1-D parameters: unfactored second moment, elementwise
//   KIND_SMALL R*C <= 64 (conv kernels [O, I, 3, 3] -> Bt = O*I matrices of 3x3; 1x1 convs): one thread per matrix
//   KIND_MAT   Bt == 1 large matrices (linear weights): 64 x 128 tiles, row/column sums by atomics into a scratch
// The same block -> (tensor, tile) partition is used by all passes; blocks find their tensor by binary search in the
// cumulative block table.
#include "common.cuh"

namespace nk {
namespace {

enum { KIND_VEC = 0, KIND_SMALL = 1, KIND_MAT = 2 };
constexpr int TILE_R = 64, TILE_C = 128, VEC_CHUNK = 4096, SMALL_PER_BLOCK = 1024;

struct AdafTensor {  // 88 bytes; mirrored by neurosis_b200/optim.py (_TENSOR_DTYPE)
    float* p;
    const float* g;
    float* vr;        // factored: row moments [Bt*R]; KIND_VEC: full second moment [n]
    float* vc;        // factored: column moments [Bt*C]
    float* exp_avg;   // first moment or null
    bf16* mirror;     // bf16 copy of p to refresh, or null
    float* scratch;   // KIND_MAT: rowacc[R] | colacc[C], zero between steps
    long long n;      // elements of the WHOLE parameter (RMS normalisation)
    int kind, Bt, R, C;
    int group, owner;  // owner: first record of the parameter this record belongs to (a [Bt, R, C] parameter with large
                       // R*C is Bt KIND_MAT records that share the p_sq / u_sq sums of their owner)
};
static_assert(sizeof(AdafTensor) == 88, "AdafTensor layout");

struct AdafHyper {  // per parameter group, refreshed by the host every step (8 floats)
    float beta2t, rel_step, eps1, eps2, clip, weight_decay, beta1, scale_parameter;
};
struct AdafScal {  // per tensor, zeroed at the start of every step
    float p_sq, u_sq, vr_mean, _pad;
};

__device__ __forceinline__ int find_tensor(const int* __restrict__ blk_start, int n_tensors, int blk) {
    int lo = 0, hi = n_tensors - 1;  // largest t with blk_start[t] <= blk
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (blk_start[mid] <= blk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

// PASS 1: p_sq, second-moment statistics.  PASS 2: sum of squared updates.  PASS 3: apply.
template <int PASS>
__global__ void __launch_bounds__(256) adafactor_kernel(const AdafTensor* __restrict__ tensors,
                                                        const int* __restrict__ blk_start, int n_tensors,
                                                        const AdafHyper* __restrict__ hyper,
                                                        AdafScal* __restrict__ scal, float* __restrict__ rms_out) {
    __shared__ float red[32];
    __shared__ float colpart[8][TILE_C];
    const int t = find_tensor(blk_start, n_tensors, blockIdx.x);
    const AdafTensor T = tensors[t];
    const AdafHyper H = hyper[T.group];
    const int lb = blockIdx.x - blk_start[t];  // block index within the tensor
    const float b2 = H.beta2t, omb2 = 1.f - H.beta2t;

    // step-size terms (PASS 3)
    float lr = 0.f, inv_clip = 1.f;
    if (PASS == 3) {
        const float n = static_cast<float>(T.n);
        const float rms_p = sqrtf(scal[T.owner].p_sq / n);
        lr = (H.scale_parameter != 0.f ? fmaxf(H.eps2, rms_p) : 1.f) * H.rel_step;
        inv_clip = 1.f / fmaxf(1.f, sqrtf(scal[T.owner].u_sq / n) / H.clip);
        if (lb == 0 && threadIdx.x == 0 && rms_out) rms_out[t] = rms_p;
    }
    auto apply = [&](long long i, float upd) {
        float pv = T.p[i];
        float d = upd * inv_clip * lr;
        if (T.exp_avg) {
            const float m = T.exp_avg[i] * H.beta1 + d * (1.f - H.beta1);
            T.exp_avg[i] = m;
            d = m;
        }
        if (H.weight_decay != 0.f) pv += pv * (-H.weight_decay * lr);
        pv -= d;
        T.p[i] = pv;
        if (T.mirror) T.mirror[i] = __float2bfloat16(pv);
    };

    float acc = 0.f;  // PASS 1: sum p^2 ; PASS 2: sum upd^2
    if (T.kind == KIND_VEC) {
        const long long i0 = static_cast<long long>(lb) * VEC_CHUNK;
        const long long i1 = i0 + VEC_CHUNK < T.n ? i0 + VEC_CHUNK : T.n;
        for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
            const float g = T.g[i];
            if (PASS == 1) {
                const float pv = T.p[i];
                acc += pv * pv;
                T.vr[i] = T.vr[i] * b2 + (g * g + H.eps1) * omb2;
            } else {
                const float upd = rsqrtf(T.vr[i]) * g;
                if (PASS == 2) acc += upd * upd;
                else apply(i, upd);
            }
        }
    } else if (T.kind == KIND_SMALL) {
        const int R = T.R, C = T.C, RC = R * C;
        const long long m0 = static_cast<long long>(lb) * SMALL_PER_BLOCK;
        for (long long m = m0 + threadIdx.x; m < m0 + SMALL_PER_BLOCK && m < T.Bt; m += blockDim.x) {
            const float* g = T.g + m * RC;
            float* vr = T.vr + m * R;
            float* vc = T.vc + m * C;
            if (PASS == 1) {
                const float* p = T.p + m * RC;
                for (int i = 0; i < RC; ++i) acc += p[i] * p[i];
                for (int r = 0; r < R; ++r) {
                    float s = 0.f;
                    for (int c = 0; c < C; ++c) s += g[r * C + c] * g[r * C + c] + H.eps1;
                    vr[r] = vr[r] * b2 + (s / static_cast<float>(C)) * omb2;
                }
                for (int c = 0; c < C; ++c) {
                    float s = 0.f;
                    for (int r = 0; r < R; ++r) s += g[r * C + c] * g[r * C + c] + H.eps1;
                    vc[c] = vc[c] * b2 + (s / static_cast<float>(R)) * omb2;
                }
            } else {
                float mean = 0.f;
                for (int r = 0; r < R; ++r) mean += vr[r];
                mean /= static_cast<float>(R);
                for (int r = 0; r < R; ++r) {
                    const float rf = rsqrtf(vr[r] / mean);
                    for (int c = 0; c < C; ++c) {
                        const float upd = rf * rsqrtf(vc[c]) * g[r * C + c];
                        if (PASS == 2) acc += upd * upd;
                        else apply(m * RC + r * C + c, upd);
                    }
                }
            }
        }
    } else {  // KIND_MAT
        const int R = T.R, C = T.C;
        const int tiles_c = (C + TILE_C - 1) / TILE_C;
        const int r0 = (lb / tiles_c) * TILE_R, c0 = (lb % tiles_c) * TILE_C;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int c = c0 + 4 * lane;
        const bool vec4 = (C % 4 == 0) && c + 3 < C && ((reinterpret_cast<uintptr_t>(T.g) & 15u) == 0) &&
                          ((reinterpret_cast<uintptr_t>(T.p) & 15u) == 0);
        float* rowacc = T.scratch;
        float* colacc = T.scratch + R;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
        float cf[4] = {0.f, 0.f, 0.f, 0.f};
        float vr_mean = 1.f;
        if (PASS != 1) {
            vr_mean = scal[t].vr_mean;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + j < C) cf[j] = rsqrtf(T.vc[c + j]);
        }
        for (int rr = warp; rr < TILE_R; rr += 8) {
            const int r = r0 + rr;
            if (r >= R) break;  // warp-uniform
            const long long base = static_cast<long long>(r) * C + c;
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            if (vec4) {
                const float4 f = *reinterpret_cast<const float4*>(T.g + base);
                g[0] = f.x; g[1] = f.y; g[2] = f.z; g[3] = f.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < C) g[j] = T.g[base + j];
            }
            if (PASS == 1) {
                float pv[4] = {0.f, 0.f, 0.f, 0.f};
                if (vec4) {
                    const float4 f = *reinterpret_cast<const float4*>(T.p + base);
                    pv[0] = f.x; pv[1] = f.y; pv[2] = f.z; pv[3] = f.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c + j < C) pv[j] = T.p[base + j];
                }
                float rs = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc += pv[j] * pv[j];
                    if (c + j < C) {
                        const float u = g[j] * g[j] + H.eps1;
                        rs += u;
                        cs[j] += u;
                    }
                }
                rs = warp_sum(rs);
                if (lane == 0) atomicAdd(rowacc + r, rs);
            } else {
                const float rf = rsqrtf(T.vr[r] / vr_mean);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (c + j < C) {
                        const float upd = rf * cf[j] * g[j];
                        if (PASS == 2) acc += upd * upd;
                        else apply(base + j, upd);
                    }
                }
            }
        }
        if (PASS == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) colpart[warp][4 * lane + j] = cs[j];
            __syncthreads();
            if (threadIdx.x < TILE_C && c0 + threadIdx.x < C) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += colpart[w][threadIdx.x];
                atomicAdd(colacc + c0 + threadIdx.x, s);
            }
        }
    }
    if (PASS != 3) {
        const float s = block_sum(acc, red);
        if (threadIdx.x == 0 && s != 0.f) atomicAdd(PASS == 1 ? &scal[T.owner].p_sq : &scal[T.owner].u_sq, s);
    }
}

// between pass 1 and 2: fold the accumulated row / column sums of the large matrices into their moments, compute the
// mean of the row moments, and clear the scratch for the next step.  One block per tensor.
__global__ void __launch_bounds__(256) adafactor_finalize_kernel(const AdafTensor* __restrict__ tensors, int n_tensors,
                                                                 const AdafHyper* __restrict__ hyper,
                                                                 AdafScal* __restrict__ scal) {
    __shared__ float red[32];
    const int t = blockIdx.x;
    const AdafTensor T = tensors[t];
    if (T.kind != KIND_MAT) return;
    const AdafHyper H = hyper[T.group];
    const float b2 = H.beta2t, omb2 = 1.f - H.beta2t;
    float* rowacc = T.scratch;
    float* colacc = T.scratch + T.R;
    float acc = 0.f;
    for (int r = threadIdx.x; r < T.R; r += blockDim.x) {
        const float v = T.vr[r] * b2 + (rowacc[r] / static_cast<float>(T.C)) * omb2;
        T.vr[r] = v;
        rowacc[r] = 0.f;
        acc += v;
    }
    for (int c = threadIdx.x; c < T.C; c += blockDim.x) {
        T.vc[c] = T.vc[c] * b2 + (colacc[c] / static_cast<float>(T.R)) * omb2;
        colacc[c] = 0.f;
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) scal[t].vr_mean = s / static_cast<float>(T.R);
}

// Device-side step scalars for captured graphs: a replayed graph cannot take new host scalars, so the step count lives
// in device memory, is advanced by this one-thread-per-group kernel, and the per-step hyper-parameters of
// adafactor.py:131-134, 216 are derived from it.
struct AdafGroupConst {  // 8 floats per parameter group
    float decay_rate, lr, eps1, eps2, clip, weight_decay, beta1, flags;  // flags: 1 relative_step, 2 warmup_init, 4 scale_parameter
};
__global__ void adafactor_hyper_kernel(const AdafGroupConst* __restrict__ consts, long long* __restrict__ step,
                                       AdafHyper* __restrict__ hyper, int n_groups) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const AdafGroupConst c = consts[g];
    const int flags = static_cast<int>(c.flags);
    const long long s = step[g] + 1;
    step[g] = s;
    const float sf = static_cast<float>(s);
    AdafHyper h;
    h.beta2t = 1.f - powf(sf, c.decay_rate);
    h.rel_step = (flags & 1) ? fminf((flags & 2) ? 1e-6f * sf : 1e-2f, rsqrtf(sf)) : c.lr;
    h.eps1 = c.eps1;
    h.eps2 = c.eps2;
    h.clip = c.clip;
    h.weight_decay = c.weight_decay;
    h.beta1 = c.beta1;
    h.scale_parameter = (flags & 4) ? 1.f : 0.f;
    hyper[g] = h;
}
// LitEma's decay warm-up from the device-resident update counter (ema.py:43-46): n = ++num_updates (if >= 0);
// one_minus_decay = 1 - min(decay, (1 + n) / (10 + n))
__global__ void ema_decay_kernel(float decay, int* __restrict__ num_updates, float* __restrict__ one_minus_decay) {
    float d = decay;
    if (*num_updates >= 0) {
        const int n = *num_updates + 1;
        *num_updates = n;
        d = fminf(decay, static_cast<float>(1 + n) / static_cast<float>(10 + n));
    }
    *one_minus_decay = 1.f - d;
}

// ---- EMA (LitEma.forward, modules/ema.py:40-59): shadow -= (1 - decay) * (shadow - p), one block per span -----------
struct EmaSpan {
    float* shadow;
    const float* p;
    long long n;
};
__global__ void __launch_bounds__(256) ema_update_multi_kernel(const EmaSpan* __restrict__ spans,
                                                               const float* __restrict__ one_minus_decay) {
    const EmaSpan sp = spans[blockIdx.x];
    const float w = *one_minus_decay;
    const bool vec4 = ((reinterpret_cast<uintptr_t>(sp.shadow) | reinterpret_cast<uintptr_t>(sp.p)) & 15u) == 0;
    long long i = 0;
    if (vec4) {
        const long long n4 = sp.n / 4;
        float4* s4 = reinterpret_cast<float4*>(sp.shadow);
        const float4* p4 = reinterpret_cast<const float4*>(sp.p);
        for (long long k = threadIdx.x; k < n4; k += blockDim.x) {
            float4 s = s4[k];
            const float4 p = p4[k];
            s.x -= w * (s.x - p.x);
            s.y -= w * (s.y - p.y);
            s.z -= w * (s.z - p.z);
            s.w -= w * (s.w - p.w);
            s4[k] = s;
        }
        i = n4 * 4;
    }
    for (long long k = i + threadIdx.x; k < sp.n; k += blockDim.x) sp.shadow[k] -= w * (sp.shadow[k] - sp.p[k]);
}

}  // namespace
}  // namespace nk

using namespace nk;
#define ST(s) ::nk::enter(s)

extern "C" {

int nk_adafactor_step(const void* tensors_dev, const int32_t* blk_start_dev, int n_tensors, int n_blocks,
                      const float* hyper_dev, float* scal_dev, float* rms_out, nk_stream_t stream) {
    NK_REQUIRE(n_tensors >= 0 && n_blocks >= 0 && (reinterpret_cast<uintptr_t>(tensors_dev) & 7u) == 0, NK_ERR_SHAPE,
               "adafactor_step: table");
    cudaStream_t st = ST(stream);
    if (n_tensors == 0 || n_blocks == 0) return NK_OK;
    const AdafTensor* T = static_cast<const AdafTensor*>(tensors_dev);
    const AdafHyper* H = reinterpret_cast<const AdafHyper*>(hyper_dev);
    AdafScal* S = reinterpret_cast<AdafScal*>(scal_dev);
    NK_CUDA(cudaMemsetAsync(scal_dev, 0, sizeof(AdafScal) * static_cast<size_t>(n_tensors), st));
    adafactor_kernel<1><<<n_blocks, 256, 0, st>>>(T, blk_start_dev, n_tensors, H, S, nullptr);
    adafactor_finalize_kernel<<<n_tensors, 256, 0, st>>>(T, n_tensors, H, S);
    adafactor_kernel<2><<<n_blocks, 256, 0, st>>>(T, blk_start_dev, n_tensors, H, S, nullptr);
    adafactor_kernel<3><<<n_blocks, 256, 0, st>>>(T, blk_start_dev, n_tensors, H, S, rms_out);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}

int nk_adafactor_hyper(const float* consts_dev, int64_t* step_dev, float* hyper_dev, int n_groups, nk_stream_t stream) {
    NK_REQUIRE(n_groups > 0 && n_groups <= 1024, NK_ERR_SHAPE, "adafactor_hyper: %d groups", n_groups);
    adafactor_hyper_kernel<<<(n_groups + 63) / 64, 64, 0, ST(stream)>>>(
        reinterpret_cast<const AdafGroupConst*>(consts_dev), reinterpret_cast<long long*>(step_dev),
        reinterpret_cast<AdafHyper*>(hyper_dev), n_groups);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}

int nk_ema_decay(float decay, int32_t* num_updates_dev, float* one_minus_decay_dev, nk_stream_t stream) {
    ema_decay_kernel<<<1, 1, 0, ST(stream)>>>(decay, num_updates_dev, one_minus_decay_dev);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}

int nk_ema_update_multi(const void* spans_dev, int n_spans, const float* one_minus_decay_dev, nk_stream_t stream) {
    NK_REQUIRE(n_spans >= 0 && (reinterpret_cast<uintptr_t>(spans_dev) & 7u) == 0, NK_ERR_SHAPE, "ema_update: span table");
    cudaStream_t st = ST(stream);
    if (n_spans == 0) return NK_OK;
    ema_update_multi_kernel<<<n_spans, 256, 0, st>>>(static_cast<const EmaSpan*>(spans_dev), one_minus_decay_dev);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}

}  // extern "C"
