// Bandwidth-bound helper kernels of the diffusion training step: GEGLU, SiLU, adds, layout
// conversion, nearest upsampling, channel concat/split, im2col for the strided / thin convs,
// column sums (bias gradients), weight packing, row softmax.  All 128-bit vectorised where the
// layout allows; fp32 math, bf16 storage.
//
// Reference call sites (under /root/reference/src/neurosis):
//   GEGLU                       modules/attention.py:50-57 (x * F.gelu(gate), exact erf GELU)
//   nn.SiLU on embeddings       modules/diffusion/openaimodel.py:273-279,586-590
//   F.interpolate nearest 2x    modules/diffusion/openaimodel.py:140
//   torch.cat skip concat       modules/diffusion/openaimodel.py:836
//   rearrange NCHW<->(HW)C      modules/attention.py:655,664
//   timestep_embedding          modules/diffusion/util.py:152-177
//   Downsample conv s2          modules/diffusion/openaimodel.py:183-190 ; model.py:65-82 (pad (0,1,0,1))
#include "common.cuh"

namespace nk {
namespace {

__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
    }
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]);
    q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = q;
}

// gelu_erf / dgelu_erf live in common.cuh (shared with the GEGLU epilogues of the GEMM kernel)

// ---- GEGLU ----------------------------------------------------------------------------------
// h: [M, 2D] = (value | gate); out[m, d] = value * gelu(gate)
__global__ void geglu_fwd_kernel(const bf16* __restrict__ h, long long ldh, bf16* __restrict__ out, long long ldo,
                                 long long M, int D) {
    const int V = D / 8;
    const long long total = M * V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long m = i / V;
        const int v = static_cast<int>(i - m * V);
        float a[8], g[8];
        ld8(h + m * ldh + v * 8, a);
        ld8(h + m * ldh + D + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] *= gelu_erf(g[j]);
        st8(out + m * ldo + v * 8, a);
    }
}
// dh[m, :D] = dout * gelu(gate);  dh[m, D:] = dout * value * gelu'(gate)
__global__ void geglu_bwd_kernel(const bf16* __restrict__ h, long long ldh, const bf16* __restrict__ dout,
                                 long long ldo, bf16* __restrict__ dh, long long lddh, long long M, int D) {
    const int V = D / 8;
    const long long total = M * V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long m = i / V;
        const int v = static_cast<int>(i - m * V);
        float a[8], g[8], d[8], da[8], dg[8];
        ld8(h + m * ldh + v * 8, a);
        ld8(h + m * ldh + D + v * 8, g);
        ld8(dout + m * ldo + v * 8, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            da[j] = d[j] * gelu_erf(g[j]);
            dg[j] = d[j] * a[j] * dgelu_erf(g[j]);
        }
        st8(dh + m * lddh + v * 8, da);
        st8(dh + m * lddh + D + v * 8, dg);
    }
}

// ---- small elementwise ------------------------------------------------------------------------
enum { EW_SILU = 0, EW_SILU_BWD = 1, EW_ADD = 2, EW_SCALE_ADD = 3 };
// y = op(a, b): SILU: silu(a); SILU_BWD: b * silu'(a); ADD: a + b; SCALE_ADD: a + alpha*b
__global__ void ew_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y, long long n,
                          int op, float alpha) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float x = __bfloat162float(a[i]);
        float r;
        if (op == EW_SILU) {
            r = x / (1.f + __expf(-x));
        } else if (op == EW_SILU_BWD) {
            const float s = 1.f / (1.f + __expf(-x));
            r = __bfloat162float(b[i]) * s * (1.f + x * (1.f - s));
        } else if (op == EW_ADD) {
            r = x + __bfloat162float(b[i]);
        } else {
            r = x + alpha * __bfloat162float(b[i]);
        }
        y[i] = __float2bfloat16(r);
    }
}
__global__ void ew_vec_add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y,
                                  long long nvec) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float x[8], z[8];
        ld8(a + i * 8, x);
        ld8(b + i * 8, z);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += z[j];
        st8(y + i * 8, x);
    }
}

// 8 elements per thread (two 16-byte loads, one 16-byte store) when both pointers are 16-byte aligned
__device__ __forceinline__ void cast8(const float* __restrict__ x, bf16* __restrict__ y) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 1);
    uint4 q;
    q.x = pack_bf16x2(a.x, a.y);
    q.y = pack_bf16x2(a.z, a.w);
    q.z = pack_bf16x2(b.x, b.y);
    q.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(y) = q;
}
__device__ __forceinline__ void cast_span(const float* __restrict__ x, bf16* __restrict__ y, long long n, long long tid,
                                          long long nthreads) {
    const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
    const long long n8 = vec ? (n >> 3) : 0;
    for (long long i = tid; i < n8; i += nthreads) cast8(x + 8 * i, y + 8 * i);
    for (long long i = 8 * n8 + tid; i < n; i += nthreads) y[i] = __float2bfloat16(x[i]);
}
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
    cast_span(x, y, n, blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x,
              static_cast<long long>(gridDim.x) * blockDim.x);
}
__global__ void cast_f32_bf16_rows_kernel(const float* __restrict__ x, long long ldx, bf16* __restrict__ y,
                                          long long ldy, long long rows, int V) {
    const long long total = rows * V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / V;
        const int v = static_cast<int>(i - r * V);
        cast8(x + r * ldx + 8 * v, y + r * ldy + 8 * v);
    }
}
// many tensors, one launch: block b converts spans[b] = {src, dst, n} (the host cuts every tensor into spans)
struct CastSpan {
    const float* src;
    bf16* dst;
    long long n;
};
__global__ void cast_f32_bf16_multi_kernel(const CastSpan* __restrict__ spans) {
    const CastSpan sp = spans[blockIdx.x];
    cast_span(sp.src, sp.dst, sp.n, threadIdx.x, blockDim.x);
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, long long n, int accumulate) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float v = __bfloat162float(x[i]);
        y[i] = accumulate ? y[i] + v : v;
    }
}

// ---- 3x3 patches of a thin NCHW fp32 image (the RGB input of the VAE encoder) ------------------------------------
// col[p, tap*C + c] = x[n, c, y+ky-1, x+kx-1] (zero outside the image), k = 9*C <= 64 values per pixel, zero-padded to
// 64: the first convolution (3 -> 128 channels) then is ONE 64-deep GEMM k-iteration per tile instead of nine taps
// over an input padded to 64 channels.  One thread per pixel: loads are coalesced along x for each (c, tap), the
// thread writes its pixel's full 128-byte row.
template <int C>
__global__ void image_patches3x3_kernel(const float* __restrict__ x, bf16* __restrict__ col, int nimg, int H, int W) {
    const long long total = static_cast<long long>(nimg) * H * W;
    for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
         p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int xw = static_cast<int>(p % W);
        const int yh = static_cast<int>((p / W) % H);
        const long long n = p / (static_cast<long long>(W) * H);
        const float* img = x + n * C * H * W;
        float v[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = yh + tap / 3 - 1, xx = xw + tap % 3 - 1;
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (in) v[tap * C + c] = __ldg(img + (static_cast<long long>(c) * H + yy) * W + xx);
        }
        uint4* dst = reinterpret_cast<uint4*>(col + p * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint4 q;
            q.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
            q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            dst[j] = q;
        }
    }
}

// ---- channel-slice copy (concat / split of NHWC tensors) -------------------------------------
__global__ void copy_channels_kernel(const bf16* __restrict__ src, long long src_stride, bf16* __restrict__ dst,
                                     long long dst_stride, long long npix, int C) {
    const int V = C / 8;
    const long long total = npix * V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long p = i / V;
        const int v = static_cast<int>(i - p * V);
        *reinterpret_cast<uint4*>(dst + p * dst_stride + v * 8) =
            *reinterpret_cast<const uint4*>(src + p * src_stride + v * 8);
    }
}

// ---- nearest 2x upsample ------------------------------------------------------------------------
__global__ void upsample2x_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int nimg, int H, int W, int C) {
    const int V = C / 8;
    const long long total = static_cast<long long>(nimg) * (2 * H) * (2 * W) * V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i % V);
        long long p = i / V;
        const int ow = static_cast<int>(p % (2 * W));
        p /= (2 * W);
        const int oh = static_cast<int>(p % (2 * H));
        const int n = static_cast<int>(p / (2 * H));
        const long long src = ((static_cast<long long>(n) * H + (oh >> 1)) * W + (ow >> 1)) * C + v * 8;
        *reinterpret_cast<uint4*>(y + i * 8) = *reinterpret_cast<const uint4*>(x + src);
    }
}
__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int nimg, int H, int W, int C) {
    const int V = C / 8;
    const long long total = static_cast<long long>(nimg) * H * W * V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i % V);
        long long p = i / V;
        const int w = static_cast<int>(p % W);
        p /= W;
        const int h = static_cast<int>(p % H);
        const int n = static_cast<int>(p / H);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float f[8];
                ld8(dy + ((static_cast<long long>(n) * 2 * H + 2 * h + a) * (2 * W) + 2 * w + b) * C + v * 8, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
        st8(dx + i * 8, acc);
    }
}

// ---- layout conversion --------------------------------------------------------------------------
// src NCHW (fp32 or bf16), dst NHWC bf16 with Cpad >= C channels (extra channels zero), value scaled
// per image by scale[n] (nullptr = 1)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ src, bf16* __restrict__ dst, const float* __restrict__ scale,
                                    int nimg, int C, int HW, int Cpad) {
    const long long total = static_cast<long long>(nimg) * HW * Cpad;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % Cpad);
        const long long p = i / Cpad;
        const int hw = static_cast<int>(p % HW);
        const int n = static_cast<int>(p / HW);
        float v = 0.f;
        if (c < C) {
            v = static_cast<float>(src[(static_cast<long long>(n) * C + c) * HW + hw]);
            if (scale) v *= scale[n];
        }
        dst[i] = __float2bfloat16(v);
    }
}
// src NHWC bf16 (pixel stride), dst NCHW (fp32 or bf16), first C channels
template <typename T>
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ src, long long src_stride, T* __restrict__ dst, int nimg,
                                    int C, int HW) {
    const long long total = static_cast<long long>(nimg) * C * HW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int hw = static_cast<int>(i % HW);
        const long long q = i / HW;
        const int c = static_cast<int>(q % C);
        const int n = static_cast<int>(q / C);
        dst[i] = static_cast<T>(__bfloat162float(src[(static_cast<long long>(n) * HW + hw) * src_stride + c]));
    }
}

// ---- im2col / col2im for strided convs ------------------------------------------------------------
// col[(n,oh,ow), tap*C + c] = x[n, oh*stride + ky - pad_t, ow*stride + kx - pad_l, c] (0 outside)
__global__ void im2col_kernel(const bf16* __restrict__ x, long long x_stride, bf16* __restrict__ col, int nimg, int H,
                              int W, int C, int ks, int stride, int pad_t, int pad_l, int Ho, int Wo) {
    const int V = C / 8;
    const long long total = static_cast<long long>(nimg) * Ho * Wo * ks * ks * V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i % V);
        long long q = i / V;
        const int tap = static_cast<int>(q % (ks * ks));
        q /= (ks * ks);
        const int ow = static_cast<int>(q % Wo);
        q /= Wo;
        const int oh = static_cast<int>(q % Ho);
        const int n = static_cast<int>(q / Ho);
        const int h = oh * stride + tap / ks - pad_t;
        const int w = ow * stride + tap % ks - pad_l;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (h >= 0 && h < H && w >= 0 && w < W)
            val = *reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * H + h) * W + w) * x_stride + v * 8);
        *reinterpret_cast<uint4*>(col + i * 8) = val;
    }
}
// dx[n,h,w,c] = sum over (oh,ow,tap) that read (h,w):  gather form, no atomics
__global__ void col2im_kernel(const bf16* __restrict__ dcol, bf16* __restrict__ dx, int nimg, int H, int W, int C,
                              int ks, int stride, int pad_t, int pad_l, int Ho, int Wo) {
    const int V = C / 8;
    const long long total = static_cast<long long>(nimg) * H * W * V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i % V);
        long long q = i / V;
        const int w = static_cast<int>(q % W);
        q /= W;
        const int h = static_cast<int>(q % H);
        const int n = static_cast<int>(q / H);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int ky = 0; ky < ks; ++ky) {
            const int th = h + pad_t - ky;
            if (th < 0 || th % stride != 0) continue;
            const int oh = th / stride;
            if (oh >= Ho) continue;
            for (int kx = 0; kx < ks; ++kx) {
                const int tw = w + pad_l - kx;
                if (tw < 0 || tw % stride != 0) continue;
                const int ow = tw / stride;
                if (ow >= Wo) continue;
                float f[8];
                ld8(dcol + (((static_cast<long long>(n) * Ho + oh) * Wo + ow) * (ks * ks) + ky * ks + kx) * C + v * 8, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
        }
        st8(dx + i * 8, acc);
    }
}

// ---- column sums: out[g, c] += sum_{r < rows_per_group} x[g*rows_per_group + r, c] ----------------------
// block = 32 column octets (256 columns, 16-byte loads) x 8 row lanes; 4 rows in flight per thread; partial sums are
// combined in shared memory and added to `out` with one atomic per column and block (out must be initialised).
__global__ void colsum_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ out, int rows_per_group,
                              int C, int row_chunks) {
    const int g = blockIdx.z;
    const int cv = blockIdx.x * 32 + (threadIdx.x & 31);  // column octet: columns 8*cv .. 8*cv+7
    const int rl = threadIdx.x >> 5;                      // 0..7
    const int rows_per_chunk = (rows_per_group + row_chunks - 1) / row_chunks;
    const int r0 = blockIdx.y * rows_per_chunk;
    const int r1 = min(rows_per_group, r0 + rows_per_chunk);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const bool col_ok = 8 * cv < C;
    if (col_ok) {
        const bf16* base = x + (static_cast<long long>(g) * rows_per_group) * ldx + 8 * cv;
        int r = r0 + rl;
        for (; r + 24 < r1; r += 32) {  // 4 independent 16-byte loads in flight
            float f0[8], f1[8], f2[8], f3[8];
            ld8(base + static_cast<long long>(r) * ldx, f0);
            ld8(base + static_cast<long long>(r + 8) * ldx, f1);
            ld8(base + static_cast<long long>(r + 16) * ldx, f2);
            ld8(base + static_cast<long long>(r + 24) * ldx, f3);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
        }
        for (; r < r1; r += 8) {
            float f0[8];
            ld8(base + static_cast<long long>(r) * ldx, f0);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f0[j];
        }
    }
    __shared__ float sm[8][32][9];
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[rl][threadIdx.x & 31][j] = acc[j];
    __syncthreads();
    // 256 threads <-> 256 columns of this block
    const int c_local = threadIdx.x;  // 0..255
    const int c = blockIdx.x * 256 + c_local;
    if (c < C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][c_local >> 3][c_local & 7];
        atomicAdd(out + static_cast<long long>(g) * C + c, t);
    }
}

// ---- sinusoidal timestep embedding: out[b, :] = cos(t*f) | sin(t*f), f_i = exp(-ln(P) i/half) -------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, bf16* __restrict__ out, int B, int dim,
                                          float max_period) {
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, k = i - b * half;
    const float freq = expf(-logf(max_period) * static_cast<float>(k) / static_cast<float>(half));
    const float arg = t[b] * freq;
    out[static_cast<long long>(b) * dim + k] = __float2bfloat16(cosf(arg));
    out[static_cast<long long>(b) * dim + half + k] = __float2bfloat16(sinf(arg));
    if ((dim & 1) && k == 0) out[static_cast<long long>(b) * dim + dim - 1] = __float2bfloat16(0.f);
}

// ---- conv weight packing ------------------------------------------------------------------------------
// w: fp32 OIHW [Co, Ci, ks, ks].
// fwd  : wp[co, tap*CiP + ci]            (co < CoP rows, zero padded)      tap = ky*ks + kx
// dgrad: wd[ci, tapf*CoP + co]           (ci < CiP rows, zero padded)      tapf = (ks-1-ky)*ks + (ks-1-kx)
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, bf16* __restrict__ wp, bf16* __restrict__ wd,
                                        int Co, int Ci, int ks, int CoP, int CiP) {
    const int taps = ks * ks;
    const long long total = static_cast<long long>(CoP) * taps * CiP;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % CiP);
        long long q = i / CiP;
        const int tap = static_cast<int>(q % taps);
        const int co = static_cast<int>(q / taps);
        float v = 0.f;
        if (co < Co && ci < Ci) v = w[((static_cast<long long>(co) * Ci + ci) * taps) + tap];
        const bf16 bv = __float2bfloat16(v);
        if (wp) wp[i] = bv;
        if (wd) wd[(static_cast<long long>(ci) * taps + (taps - 1 - tap)) * CoP + co] = bv;
    }
}
// Same packing, tiled through shared memory so that all three global streams are coalesced: a block owns 32 output
// x 32 input channels, reads their (32 x 32*taps) fp32 block as contiguous row segments, and writes 64-byte bf16 row
// segments of both packings (the scalar kernel above reads with a stride of `taps` floats and scatters 2-byte stores
// for the transposed packing).  Requires Co, Ci, CoP, CiP multiples of 32.
__global__ void __launch_bounds__(256) pack_conv_weight_tiled_kernel(const float* __restrict__ w, bf16* __restrict__ wp,
                                                                     bf16* __restrict__ wd, int Co, int Ci, int taps,
                                                                     int CoP, int CiP) {
    extern __shared__ bf16 tile[];  // [32 co][32*taps + 2]  (ci-major within a row: index ci*taps + tap)
    const int ld = 32 * taps + 2;
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int row_len = 32 * taps;  // floats per co row of this block, contiguous in w
    for (int idx = threadIdx.x; idx < 32 * row_len; idx += blockDim.x) {
        const int r = idx / row_len, c = idx - r * row_len;
        const int co = co0 + r;
        const int ci = ci0 + c / taps;
        float v = 0.f;
        if (co < Co && ci < Ci) v = w[(static_cast<long long>(co) * Ci + ci0) * taps + c];
        tile[r * ld + c] = __float2bfloat16(v);
    }
    __syncthreads();
    // forward packing: wp[co, tap*CiP + ci]  -> for each (co, tap): 32 consecutive ci
    if (wp) {
        for (int idx = threadIdx.x; idx < 32 * taps * 32; idx += blockDim.x) {
            const int ci = idx & 31;
            const int t = (idx >> 5) % taps;
            const int r = idx / (32 * taps);
            wp[(static_cast<long long>(co0 + r) * taps + t) * CiP + ci0 + ci] = tile[r * ld + ci * taps + t];
        }
    }
    // data-gradient packing: wd[ci, tapf*CoP + co], tapf = taps-1-tap -> for each (ci, tap): 32 consecutive co
    if (wd) {
        for (int idx = threadIdx.x; idx < 32 * taps * 32; idx += blockDim.x) {
            const int r = idx & 31;  // co within the block (fastest: contiguous in wd)
            const int t = (idx >> 5) % taps;
            const int ci = idx / (32 * taps);
            wd[(static_cast<long long>(ci0 + ci) * taps + (taps - 1 - t)) * CoP + co0 + r] = tile[r * ld + ci * taps + t];
        }
    }
}
// dw[co, ci, ky, kx] += dwp[co, tap, ci]   (dwp row stride = taps*CiP)
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Co, int Ci, int ks,
                                         int CiP, int accumulate) {
    const int taps = ks * ks;
    const long long total = static_cast<long long>(Co) * Ci * taps;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i % taps);
        long long q = i / taps;
        const int ci = static_cast<int>(q % Ci);
        const int co = static_cast<int>(q / Ci);
        const float v = dwp[(static_cast<long long>(co) * taps + tap) * CiP + ci];
        dw[i] = accumulate ? dw[i] + v : v;
    }
}

// ---- row softmax for the materialised attention path -------------------------------------------------
// S fp32 [rows, ld] -> P bf16 [rows, ldp] = softmax(scale * S[:, :N]) ; lse[row] = log-sum-exp (natural log)
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
    v = is_max ? warp_max(v) : warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : (is_max ? -INFINITY : 0.f);
        t = is_max ? warp_max(t) : warp_sum(t);
        if (threadIdx.x == 0) red[32] = t;
    }
    __syncthreads();
    const float out = red[32];
    __syncthreads();
    return out;
}

// one block per row; the row (<= 256 threads x VPT float4) is read from global ONCE and kept in registers
template <int VPT>
__global__ void softmax_rows_reg_kernel(const float* __restrict__ S, long long lds, bf16* __restrict__ P, long long ldp,
                                        float* __restrict__ lse, int N, float scale) {
    __shared__ float red[33];
    const long long row = blockIdx.x;
    const float4* s4 = reinterpret_cast<const float4*>(S + row * lds);
    const int n4 = N >> 2;
    float4 v[VPT];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int idx = threadIdx.x + i * blockDim.x;
        if (idx < n4) {
            v[i] = __ldg(s4 + idx);
            mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
        }
    }
    mx = block_reduce(mx, true, red);
    const float sl2 = scale * 1.4426950408889634f;
    const float mb = mx * sl2;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int idx = threadIdx.x + i * blockDim.x;
        if (idx < n4) {
            v[i].x = exp2f(fmaf(v[i].x, sl2, -mb));
            v[i].y = exp2f(fmaf(v[i].y, sl2, -mb));
            v[i].z = exp2f(fmaf(v[i].z, sl2, -mb));
            v[i].w = exp2f(fmaf(v[i].w, sl2, -mb));
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    sum = block_reduce(sum, false, red);
    const float inv = 1.f / sum;
    uint2* p2 = reinterpret_cast<uint2*>(P + row * ldp);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int idx = threadIdx.x + i * blockDim.x;
        if (idx < n4) p2[idx] = make_uint2(pack_bf16x2(v[i].x * inv, v[i].y * inv), pack_bf16x2(v[i].z * inv, v[i].w * inv));
    }
    if (threadIdx.x == 0 && lse) lse[row] = mx * scale + logf(sum);
}

__global__ void softmax_rows_kernel(const float* __restrict__ S, long long lds, bf16* __restrict__ P, long long ldp,
                                    float* __restrict__ lse, long long rows, int N, float scale) {
    __shared__ float red[33];
    const long long row = blockIdx.x;
    if (row >= rows) return;
    const float* s = S + row * lds;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < N; i += blockDim.x) mx = fmaxf(mx, s[i]);
    mx = block_reduce(mx, true, red);
    float sum = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) sum += __expf((s[i] - mx) * scale);
    sum = block_reduce(sum, false, red);
    const float inv = 1.f / sum;
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        P[row * ldp + i] = __float2bfloat16(__expf((s[i] - mx) * scale) * inv);
    if (threadIdx.x == 0 && lse) lse[row] = mx * scale + logf(sum);
}

// delta[b, h, q] = sum_d dO[b,q,h,d] * O[b,q,h,d]     (tensors laid out [B, N, H, D])
__global__ void attn_delta_kernel(const bf16* __restrict__ dO, const bf16* __restrict__ O, float* __restrict__ delta,
                                  int B, int N, int H, int D) {
    const long long total = static_cast<long long>(B) * N * H;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x >> 3) + (threadIdx.x >> 3);  // 8 lanes per row
    const int sub = threadIdx.x & 7;
    float acc = 0.f;
    if (idx < total) {
        const bf16* a = dO + idx * D;
        const bf16* b = O + idx * D;
        for (int d = sub * 8; d < D; d += 64) {
            float x[8], y[8];
            ld8(a + d, x);
            ld8(b + d, y);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += x[j] * y[j];
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (idx < total && sub == 0) {
        const int h = static_cast<int>(idx % H);
        const long long bn = idx / H;
        const int q = static_cast<int>(bn % N);
        const int b = static_cast<int>(bn / N);
        delta[(static_cast<long long>(b) * H + h) * N + q] = acc;
    }
}

inline int grid_for(long long work, int block) {
    const long long g = (work + block - 1) / block;
    return static_cast<int>(std::max<long long>(1, std::min<long long>(g, 148LL * 16)));
}

}  // namespace
}  // namespace nk

using namespace nk;
#define ST(s) ::nk::enter(s)
#define BF(p) static_cast<bf16*>(p)
#define CBF(p) static_cast<const bf16*>(p)

extern "C" {

int nk_geglu_fwd(const void* h, int64_t ldh, void* out, int64_t ldo, int64_t M, int D, nk_stream_t stream) {
    NK_REQUIRE(D % 8 == 0 && ldh % 8 == 0 && ldo % 8 == 0, NK_ERR_SHAPE, "geglu: D=%d", D);
    geglu_fwd_kernel<<<grid_for(M * (D / 8), 256), 256, 0, ST(stream)>>>(CBF(h), ldh, BF(out), ldo, M, D);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_geglu_bwd(const void* h, int64_t ldh, const void* dout, int64_t ldo, void* dh, int64_t lddh, int64_t M, int D,
                 nk_stream_t stream) {
    NK_REQUIRE(D % 8 == 0 && ldh % 8 == 0 && ldo % 8 == 0 && lddh % 8 == 0, NK_ERR_SHAPE, "geglu: D=%d", D);
    geglu_bwd_kernel<<<grid_for(M * (D / 8), 256), 256, 0, ST(stream)>>>(CBF(h), ldh, CBF(dout), ldo, BF(dh), lddh, M, D);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_silu_fwd(const void* x, void* y, int64_t n, nk_stream_t stream) {
    ew_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(CBF(x), nullptr, BF(y), n, EW_SILU, 0.f);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, nk_stream_t stream) {
    ew_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(CBF(x), CBF(dy), BF(dx), n, EW_SILU_BWD, 0.f);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_add(const void* a, const void* b, void* y, int64_t n, nk_stream_t stream) {
    if (n % 8 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0)
        ew_vec_add_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>(CBF(a), CBF(b), BF(y), n / 8);
    else
        ew_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(CBF(a), CBF(b), BF(y), n, EW_ADD, 1.f);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_cast_f32_bf16(const float* x, void* y, int64_t n, nk_stream_t stream) {
    cast_f32_bf16_kernel<<<grid_for((n + 7) / 8, 256), 256, 0, ST(stream)>>>(x, BF(y), n);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_cast_f32_bf16_rows(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int cols,
                          nk_stream_t stream) {
    NK_REQUIRE(cols % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(y) & 15u) == 0,
               NK_ERR_SHAPE, "cast rows: cols/ld must be multiples of 8 and pointers 16-byte aligned");
    cudaStream_t st = ST(stream);
    if (rows <= 0 || cols <= 0) return NK_OK;
    cast_f32_bf16_rows_kernel<<<grid_for(rows * (cols / 8), 256), 256, 0, st>>>(x, ldx, BF(y), ldy, rows, cols / 8);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_cast_f32_bf16_multi(const void* spans_dev, int n_spans, nk_stream_t stream) {
    NK_REQUIRE(n_spans >= 0 && (reinterpret_cast<uintptr_t>(spans_dev) & 7u) == 0, NK_ERR_SHAPE, "cast_multi: span table");
    cudaStream_t st = ST(stream);
    if (n_spans == 0) return NK_OK;
    cast_f32_bf16_multi_kernel<<<n_spans, 256, 0, st>>>(static_cast<const CastSpan*>(spans_dev));
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_cast_bf16_f32(const void* x, float* y, int64_t n, int accumulate, nk_stream_t stream) {
    cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(CBF(x), y, n, accumulate);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_image_patches3x3(const float* x, void* col, int nimg, int C, int H, int W, nk_stream_t stream) {
    NK_REQUIRE((C == 1 || C == 3 || C == 4) && nimg > 0 && H > 0 && W > 0, NK_ERR_SHAPE,
               "image_patches3x3: C=%d (1, 3 or 4)", C);
    cudaStream_t st = ST(stream);
    const int grid = grid_for(1LL * nimg * H * W, 128);
    if (C == 3) image_patches3x3_kernel<3><<<grid, 128, 0, st>>>(x, BF(col), nimg, H, W);
    else if (C == 4) image_patches3x3_kernel<4><<<grid, 128, 0, st>>>(x, BF(col), nimg, H, W);
    else image_patches3x3_kernel<1><<<grid, 128, 0, st>>>(x, BF(col), nimg, H, W);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_copy_channels(const void* src, int64_t src_stride, void* dst, int64_t dst_stride, int64_t npix, int C,
                     nk_stream_t stream) {
    NK_REQUIRE(C % 8 == 0 && src_stride % 8 == 0 && dst_stride % 8 == 0, NK_ERR_SHAPE, "copy_channels: C=%d", C);
    copy_channels_kernel<<<grid_for(npix * (C / 8), 256), 256, 0, ST(stream)>>>(CBF(src), src_stride, BF(dst), dst_stride, npix, C);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_upsample2x_fwd(const void* x, void* y, int nimg, int H, int W, int C, nk_stream_t stream) {
    NK_REQUIRE(C % 8 == 0, NK_ERR_SHAPE, "upsample: C=%d", C);
    upsample2x_fwd_kernel<<<grid_for(4LL * nimg * H * W * (C / 8), 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), nimg, H, W, C);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_upsample2x_bwd(const void* dy, void* dx, int nimg, int H, int W, int C, nk_stream_t stream) {
    NK_REQUIRE(C % 8 == 0, NK_ERR_SHAPE, "upsample: C=%d", C);
    upsample2x_bwd_kernel<<<grid_for(1LL * nimg * H * W * (C / 8), 256), 256, 0, ST(stream)>>>(CBF(dy), BF(dx), nimg, H, W, C);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_nchw_to_nhwc(const void* src, int src_is_f32, void* dst, const float* scale, int nimg, int C, int HW, int Cpad,
                    nk_stream_t stream) {
    const long long total = 1LL * nimg * HW * Cpad;
    if (src_is_f32)
        nchw_to_nhwc_kernel<float><<<grid_for(total, 256), 256, 0, ST(stream)>>>(static_cast<const float*>(src), BF(dst), scale, nimg, C, HW, Cpad);
    else
        nchw_to_nhwc_kernel<bf16><<<grid_for(total, 256), 256, 0, ST(stream)>>>(CBF(src), BF(dst), scale, nimg, C, HW, Cpad);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_nhwc_to_nchw(const void* src, int64_t src_stride, void* dst, int dst_is_f32, int nimg, int C, int HW,
                    nk_stream_t stream) {
    const long long total = 1LL * nimg * HW * C;
    if (dst_is_f32)
        nhwc_to_nchw_kernel<float><<<grid_for(total, 256), 256, 0, ST(stream)>>>(CBF(src), src_stride, static_cast<float*>(dst), nimg, C, HW);
    else
        nhwc_to_nchw_kernel<bf16><<<grid_for(total, 256), 256, 0, ST(stream)>>>(CBF(src), src_stride, BF(dst), nimg, C, HW);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_im2col(const void* x, int64_t x_stride, void* col, int nimg, int H, int W, int C, int ks, int stride, int pad_t,
              int pad_l, int Ho, int Wo, nk_stream_t stream) {
    NK_REQUIRE(C % 8 == 0 && x_stride % 8 == 0, NK_ERR_SHAPE, "im2col: C=%d", C);
    im2col_kernel<<<grid_for(1LL * nimg * Ho * Wo * ks * ks * (C / 8), 256), 256, 0, ST(stream)>>>(
        CBF(x), x_stride, BF(col), nimg, H, W, C, ks, stride, pad_t, pad_l, Ho, Wo);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_col2im(const void* dcol, void* dx, int nimg, int H, int W, int C, int ks, int stride, int pad_t, int pad_l,
              int Ho, int Wo, nk_stream_t stream) {
    NK_REQUIRE(C % 8 == 0, NK_ERR_SHAPE, "col2im: C=%d", C);
    col2im_kernel<<<grid_for(1LL * nimg * H * W * (C / 8), 256), 256, 0, ST(stream)>>>(CBF(dcol), BF(dx), nimg, H, W, C, ks,
                                                                                    stride, pad_t, pad_l, Ho, Wo);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_colsum(const void* x, int64_t ldx, float* out, int groups, int rows_per_group, int C, int accumulate,
              nk_stream_t stream) {
    NK_REQUIRE(C % 8 == 0 && ldx % 8 == 0, NK_ERR_SHAPE, "colsum: C=%d ldx=%lld must be multiples of 8", C,
               static_cast<long long>(ldx));
    if (!accumulate)
        NK_CUDA(cudaMemsetAsync(out, 0, static_cast<size_t>(groups) * C * sizeof(float), ST(stream)));
    const int col_blocks = (C + 255) / 256;
    // enough blocks to fill the machine, at least 64 rows each
    int row_chunks = std::max(1, std::min(rows_per_group / 64, (148 * 8) / std::max(1, col_blocks * groups)));
    colsum_kernel<<<dim3(col_blocks, row_chunks, groups), 256, 0, ST(stream)>>>(CBF(x), ldx, out, rows_per_group, C,
                                                                               row_chunks);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_timestep_embedding(const float* t, void* out, int B, int dim, float max_period, nk_stream_t stream) {
    const int n = B * (dim / 2);
    timestep_embedding_kernel<<<(n + 127) / 128, 128, 0, ST(stream)>>>(t, BF(out), B, dim, max_period);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_conv_pack_weights(const float* w, void* wp_fwd, void* wp_dgrad, int Co, int Ci, int ks, int CoP, int CiP,
                         nk_stream_t stream) {
    NK_REQUIRE(CoP >= Co && CiP >= Ci, NK_ERR_SHAPE, "pack: padded dims too small");
    cudaStream_t st = ST(stream);
    const int taps = ks * ks;
    if (Co % 32 == 0 && Ci % 32 == 0 && CoP == Co && CiP == Ci && taps <= 9) {
        const size_t smem = static_cast<size_t>(32) * (32 * taps + 2) * sizeof(bf16);
        pack_conv_weight_tiled_kernel<<<dim3(Ci / 32, Co / 32), 256, smem, st>>>(w, BF(wp_fwd), BF(wp_dgrad), Co, Ci, taps,
                                                                                CoP, CiP);
    } else {
        pack_conv_weight_kernel<<<grid_for(1LL * CoP * ks * ks * CiP, 256), 256, 0, st>>>(w, BF(wp_fwd), BF(wp_dgrad), Co,
                                                                                         Ci, ks, CoP, CiP);
    }
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_conv_unpack_wgrad(const float* dw_packed, float* dw, int Co, int Ci, int ks, int CiP, int accumulate,
                         nk_stream_t stream) {
    unpack_conv_wgrad_kernel<<<grid_for(1LL * Co * Ci * ks * ks, 256), 256, 0, ST(stream)>>>(dw_packed, dw, Co, Ci, ks, CiP,
                                                                                          accumulate);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_softmax_rows(const float* S, int64_t lds, void* P, int64_t ldp, float* lse, int64_t rows, int N, float scale,
                    nk_stream_t stream) {
    NK_REQUIRE(rows < (1LL << 31), NK_ERR_SHAPE, "softmax rows");
    const bool vec_ok = (N % 4 == 0) && (lds % 4 == 0) && (ldp % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(S) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(P) & 7u) == 0);
    const unsigned grid = static_cast<unsigned>(rows);
    if (vec_ok && N <= 256 * 4 * 4)
        softmax_rows_reg_kernel<4><<<grid, 256, 0, ST(stream)>>>(S, lds, BF(P), ldp, lse, N, scale);
    else if (vec_ok && N <= 256 * 4 * 16)
        softmax_rows_reg_kernel<16><<<grid, 256, 0, ST(stream)>>>(S, lds, BF(P), ldp, lse, N, scale);
    else
        softmax_rows_kernel<<<grid, N >= 1024 ? 256 : 128, 0, ST(stream)>>>(S, lds, BF(P), ldp, lse, rows, N, scale);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}
int nk_attn_delta(const void* dO, const void* O, float* delta, int B, int N, int H, int D, nk_stream_t stream) {
    NK_REQUIRE(D % 8 == 0, NK_ERR_SHAPE, "attn_delta: D=%d", D);
    const long long rows = 1LL * B * N * H;
    attn_delta_kernel<<<static_cast<unsigned>((rows + 31) / 32), 256, 0, ST(stream)>>>(CBF(dO), CBF(O), delta, B, N, H, D);
    NK_CUDA(cudaGetLastError());
    return NK_OK;
}

}  // extern "C"
