// Noise mixing, denoiser preconditioning and the per-sample weighted MSE of the diffusion loss.
// Tiny tensors (C*H*W = 65 536 per SDXL latent) — these kernels exist to remove a dozen
// launch-bound ATen kernels and to emit the UNet input directly in its NHWC bf16 layout.
//
// Reference (under /root/reference/src/neurosis/modules/diffusion):
//   loss.py:117-146       sigma draw, noise mix  z_t = x + sigma*n   (edm)  /  (1-sigma) x + sigma n  (rf)
//   denoiser.py:40-57     c_in scaling of the network input, D = F*c_out + z_t*c_skip (nk_lincomb_per_sample)
//   loss.py:153-155 + ../losses/functions.py:81-94   loss[b] = mean((D - target)^2) * w[b]  (fp32)
#include "common.cuh"

namespace nk {
namespace {

// z[b, i] = a_b * x[b, i] + sigma_b * noise[b, i];  a_b = 1 (edm) or 1 - sigma_b (rf)
__global__ void noise_mix_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                 const float* __restrict__ sigma, float* __restrict__ z, long long per_sample,
                                 long long total, int rf) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float s = sigma[i / per_sample];
        const float a = rf ? 1.f - s : 1.f;
        z[i] = a * x[i] + s * noise[i];
    }
}

// y[b, i] = s[b] * x[b, i]
__global__ void scale_per_sample_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y,
                                        long long per_sample, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        y[i] = x[i] * s[i / per_sample];
}

// out[b, i] = a[b] * x[b, i] + c[b] * y[b, i]   (y may be null)
__global__ void lincomb_per_sample_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                          const float* __restrict__ y, const float* __restrict__ c,
                                          float* __restrict__ out, long long per_sample, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / per_sample;
        float v = a[b] * x[i];
        if (y) v += c[b] * y[i];
        out[i] = v;
    }
}

// loss[b] = w[b] * mean_i (D[b,i] - T[b,i])^2 ; one block per sample, fixed-order reduction
__global__ void weighted_mse_fwd_kernel(const float* __restrict__ D, const float* __restrict__ T,
                                        const float* __restrict__ w, float* __restrict__ loss, long long n) {
    const int b = blockIdx.x;
    const float* d = D + b * n;
    const float* t = T + b * n;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float e = d[i] - t[i];
        acc += e * e;
    }
    __shared__ float red[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) loss[b] = v / static_cast<float>(n) * w[b];
    }
}
// dD[b,i] = dloss[b] * w[b] * 2/n * (D - T)
__global__ void weighted_mse_bwd_kernel(const float* __restrict__ D, const float* __restrict__ T,
                                        const float* __restrict__ w, const float* __restrict__ dloss,
                                        float* __restrict__ dD, long long n, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / n;
        dD[i] = dloss[b] * w[b] * (2.f / static_cast<float>(n)) * (D[i] - T[i]);
    }
}

// L1 form (BatchL1Loss, losses/functions.py:65-78; StandardDiffusionLoss loss_type "l1"): loss[b] = w[b] * mean_i |D - T|
__global__ void weighted_l1_fwd_kernel(const float* __restrict__ D, const float* __restrict__ T,
                                       const float* __restrict__ w, float* __restrict__ loss, long long n) {
    const int b = blockIdx.x;
    const float* d = D + b * n;
    const float* t = T + b * n;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += fabsf(d[i] - t[i]);
    __shared__ float red[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) loss[b] = v / static_cast<float>(n) * w[b];
    }
}
// dD[b,i] = dloss[b] * w[b] / n * sign(D - T)   (sign(0) = 0, as torch's l1_loss backward)
__global__ void weighted_l1_bwd_kernel(const float* __restrict__ D, const float* __restrict__ T,
                                       const float* __restrict__ w, const float* __restrict__ dloss,
                                       float* __restrict__ dD, long long n, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / n;
        const float e = D[i] - T[i];
        const float sg = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
        dD[i] = dloss[b] * w[b] * (1.f / static_cast<float>(n)) * sg;
    }
}

// ---- diagonal Gaussian posterior of the VAE (modules/distributions.py:29-51, regularizers.py:31-41) ----------------
// moments[b] = (mean | logvar) halves of 2*C channels, each `half` = C*H*W contiguous floats (NCHW).  logvar is clamped
// to [-30, 20]; z = mean + exp(0.5 logvar) * eps (eps == nullptr: mode, z = mean); kl[b] = 0.5 * sum(mean^2 + var - 1
// - logvar).  One block per sample, fixed-order reduction.
__global__ void diag_gaussian_fwd_kernel(const float* __restrict__ moments, const float* __restrict__ eps,
                                         float* __restrict__ z, float* __restrict__ kl, long long half) {
    const int b = blockIdx.x;
    const float* mean = moments + 2 * half * b;
    const float* logvar = mean + half;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < half; i += blockDim.x) {
        const float m = mean[i];
        const float lv = fminf(fmaxf(logvar[i], -30.f), 20.f);
        const float sd = expf(0.5f * lv);
        z[half * b + i] = eps ? m + sd * eps[half * b + i] : m;
        acc += m * m + expf(lv) - 1.f - lv;
    }
    __shared__ float red[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0 && kl) kl[b] = 0.5f * v;
    }
}
// d moments from dz (may be null) and dkl[b] (may be null): dmean = dz + dkl*mean;
// dlogvar = [clamp inactive] * (dz * eps * 0.5 * sd + dkl * 0.5 * (var - 1))
__global__ void diag_gaussian_bwd_kernel(const float* __restrict__ moments, const float* __restrict__ eps,
                                         const float* __restrict__ dz, const float* __restrict__ dkl,
                                         float* __restrict__ dmoments, long long half, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / half, j = i - b * half;
        const float m = moments[2 * half * b + j];
        const float lvr = moments[2 * half * b + half + j];
        const float lv = fminf(fmaxf(lvr, -30.f), 20.f);
        const float g = dz ? dz[i] : 0.f;
        const float gk = dkl ? dkl[b] : 0.f;
        float dlv = 0.f;
        if (lvr >= -30.f && lvr <= 20.f) {
            if (eps) dlv += g * eps[i] * 0.5f * expf(0.5f * lv);
            dlv += gk * 0.5f * (expf(lv) - 1.f);
        }
        dmoments[2 * half * b + j] = g + gk * m;
        dmoments[2 * half * b + half + j] = dlv;
    }
}

inline int grid_for(long long work, int block) {
    const long long g = (work + block - 1) / block;
    return static_cast<int>(std::max<long long>(1, std::min<long long>(g, 148LL * 8)));
}

}  // namespace
}  // namespace nk

using namespace nk;
#define ST(s) ::nk::enter(s)

extern "C" {

int nk_noise_mix(const float* x, const float* noise, const float* sigma, float* z, int B, int64_t per_sample,
                 int rectified_flow, nk_stream_t stream) {
    const long long total = static_cast<long long>(B) * per_sample;
    noise_mix_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(x, noise, sigma, z, per_sample, total, rectified_flow);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_scale_per_sample(const float* x, const float* s, float* y, int B, int64_t per_sample, nk_stream_t stream) {
    const long long total = static_cast<long long>(B) * per_sample;
    scale_per_sample_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(x, s, y, per_sample, total);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_lincomb_per_sample(const float* x, const float* a, const float* y, const float* c, float* out, int B,
                          int64_t per_sample, nk_stream_t stream) {
    const long long total = static_cast<long long>(B) * per_sample;
    lincomb_per_sample_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(x, a, y, c, out, per_sample, total);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_weighted_mse_fwd(const float* D, const float* T, const float* w, float* loss, int B, int64_t per_sample,
                        nk_stream_t stream) {
    weighted_mse_fwd_kernel<<<B, 1024, 0, ST(stream)>>>(D, T, w, loss, per_sample);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_weighted_mse_bwd(const float* D, const float* T, const float* w, const float* dloss, float* dD, int B,
                        int64_t per_sample, nk_stream_t stream) {
    const long long total = static_cast<long long>(B) * per_sample;
    weighted_mse_bwd_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(D, T, w, dloss, dD, per_sample, total);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}

int nk_weighted_l1_fwd(const float* D, const float* T, const float* w, float* loss, int B, int64_t per_sample,
                       nk_stream_t stream) {
    weighted_l1_fwd_kernel<<<B, 1024, 0, ST(stream)>>>(D, T, w, loss, per_sample);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_weighted_l1_bwd(const float* D, const float* T, const float* w, const float* dloss, float* dD, int B,
                       int64_t per_sample, nk_stream_t stream) {
    const long long total = static_cast<long long>(B) * per_sample;
    weighted_l1_bwd_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(D, T, w, dloss, dD, per_sample, total);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_diag_gaussian_fwd(const float* moments, const float* eps, float* z, float* kl, int B, int64_t half,
                         nk_stream_t stream) {
    NK_REQUIRE(B > 0 && half > 0, NK_ERR_SHAPE, "diag_gaussian_fwd: B=%d", B);
    diag_gaussian_fwd_kernel<<<B, 1024, 0, ST(stream)>>>(moments, eps, z, kl, half);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_diag_gaussian_bwd(const float* moments, const float* eps, const float* dz, const float* dkl, float* dmoments,
                         int B, int64_t half, nk_stream_t stream) {
    NK_REQUIRE(B > 0 && half > 0, NK_ERR_SHAPE, "diag_gaussian_bwd: B=%d", B);
    const long long total = static_cast<long long>(B) * half;
    diag_gaussian_bwd_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(moments, eps, dz, dkl, dmoments, half, total);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}

}  // extern "C"
