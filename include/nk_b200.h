/*
 * nk_b200.h — C ABI of the neurosis_b200 sm_100a kernel library (libnk_b200.so).
 *
 * Boundary.  The reference (neggles/neurosis) has no native/FFI layer of its own: every op on
 * its diffusion-training hot path is a PyTorch library dispatch from a Python nn.Module.  The
 * entry points below are what those dispatch sites bind to in the drop-in modules of
 * `neurosis_b200.modules` (the binding is ctypes, see neurosis_b200/_lib.py and INTEGRATION.md).
 * Each declaration cites the reference call site (path relative to /root/reference/src/neurosis)
 * it replaces.
 *
 * Conventions (all entry points):
 *   - plain pointers to DEVICE memory, explicit sizes/strides in ELEMENTS, a cudaStream_t passed
 *     as void*; nothing is allocated, nothing throws;
 *   - return 0 on success, negative on failure (NK_ERR_*); nk_last_error() gives the message;
 *   - activations are bf16, NHWC ("channels last") for image tensors, row-major [tokens, C] for
 *     token tensors; statistics, losses and weight gradients are fp32;
 *   - work is enqueued on the given stream and is asynchronous with respect to the host.
 */
#ifndef NK_B200_H_
#define NK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NK_OK 0
#define NK_ERR_SHAPE (-1)
#define NK_ERR_UNSUPPORTED (-2)
#define NK_ERR_CUDA (-3)
#define NK_ERR_WORKSPACE (-4)

typedef void* nk_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------- */
int nk_version(void);
const char* nk_last_error(void);
int nk_sm_count(void);

/* ---- tensor-core GEMM core (tcgen05 + TMEM + TMA) ------------------------------------------ */

/* One GEMM operand.  Matrix form: rows x inner (inner contiguous) with two optional batch dims.
 * Image form (conv = 1): NHWC tensor (nimg, H, W, inner) whose pixel stride is row_stride.   */
typedef struct nk_operand {
    const void* ptr;   /* bf16 */
    int32_t mn_major;  /* 0: inner is the reduction dim K; 1: inner is the M (A) / N (B) dim */
    int32_t conv;
    int64_t inner, rows, row_stride;
    int64_t nb2, b2_stride, nb1, b1_stride;
    int32_t H, W, nimg, _pad;
} nk_operand;

enum { NK_EPI_LINEAR = 0, NK_EPI_EXP2 = 1, NK_EPI_DSOFTMAX = 2 };
enum { NK_OUT_BF16 = 0, NK_OUT_F32 = 1, NK_OUT_F32_ATOMIC = 2 };

/* C[b1,b2][M,N] = epilogue( sum_k A[m,k] * B[n,k] ).  See gemm_tc.cuh for field semantics. */
typedef struct nk_gemm_desc {
    nk_operand A, B;
    int32_t M, N, K;
    int32_t nb2, nb1;
    int32_t ksize, pad; /* conv forward/dgrad: A.conv = 1 */
    int32_t wgrad;      /* conv weight gradient: A.conv = B.conv = 1, N = taps*Cin (column = tap*Cin + ci) */
    void* C;
    int64_t ldc, c_b2_stride, c_b1_stride;
    int32_t out, epi;
    float alpha;
    int32_t rows_per_img;
    const float* bias;
    const float* bias_img;
    const void* residual; /* bf16 */
    int64_t ldr;
    const float* rowvec;
    const void* aux; /* bf16 */
    int32_t force_bn, force_splits;
    int32_t force_cta_group; /* 0 = heuristic, 1 = single-CTA tiles, 2 = CTA pairs (cta_group::2) */
    int32_t force_dual;      /* row-tile pairing: 0 = library mode (nk_gemm_set_dual), -1 = off, 1 = cost model, 2 = wherever legal */
} nk_gemm_desc;

int nk_gemm_ex(const nk_gemm_desc* d, nk_stream_t stream);

/* Row-tile pairing of the tensor-core GEMM / implicit-GEMM convolution kernels (gemm_tc.cu, the DUAL instantiations: a CTA
 * owns two 128-row tiles that share one B tile, 25 % fewer operand bytes per FLOP through the L2 -> SM fabric that bounds
 * the kernel).  mode 0 = off (the library default, or NK_GEMM_DUAL), 1 = where the launch cost model expects a gain,
 * 2 = wherever legal; anything else only queries.  Returns the previous mode.  A scheduling choice only: results are
 * bit-identical to the unpaired kernels (same k order per output element) except for the fp32 accumulation order of
 * split-K weight gradients.  neurosis_b200.tune enables it after checking exactly that on the device it runs on.
 * No reference counterpart (the reference delegates every contraction to cuBLAS / cuDNN: modules/attention.py:283-290,
 * modules/diffusion/openaimodel.py:247-301). */
int nk_gemm_set_dual(int mode);
/* Mode 1 pairs only launches whose reduction has at least k_iters 64-deep k-iterations (K / 64; convolutions: taps * Cin / 64):
 * with a short reduction the two epilogues that pairing exposes per scheduler tile outweigh the saved operand traffic.
 * neurosis_b200.tune measures the break-even on the device; NK_GEMM_DUAL_MIN_K pins it.  k_iters < 0 only queries.
 * Returns the previous value (default 0 = no limit). */
int nk_gemm_set_dual_min_k(int k_iters);
/* Mode 1 only: which classes of launches may pair — bit 0 matrix GEMM with K-major A (nk_linear_fwd / nk_linear_dgrad),
 * bit 1 matrix GEMM with MN-major A (nk_linear_wgrad), bit 2 implicit-GEMM convolution (nk_conv2d_fwd, also used for the data
 * gradient, nk_conv2d_stride2_fwd).  Default 7 (or NK_GEMM_DUAL_CLASSES); neurosis_b200.tune clears the classes that lose on the
 * device.  0..7 sets, anything else queries; returns the previous mask. */
int nk_gemm_set_dual_classes(int mask);
/* Paired launches: the MMA issuer lets the second row tile trail the first by k_iters k-iterations (0 = interleaved; the
 * kernel clamps to ring depth - 1), so that the first tile starts while the epilogue still drains the other TMEM buffer
 * and reaches its own epilogue earlier.  Does not change results.  0..7 sets, anything else queries; returns the previous
 * value (default 0, or NK_GEMM_DUAL_SKEW). */
int nk_gemm_set_dual_skew(int k_iters);
/* Epilogue side-input prefetch: at the start of every output tile one thread per epilogue warp-half issues
 * `cp.async.bulk.prefetch.tensor.L2` for the boxes of the residual (EPI_LINEAR) or of the saved GEGLU pre-activation h
 * (nk_linear_dgrad_geglu) that the epilogue will read — while the tile's main loop still runs.  The epilogue reads that
 * input with one 32-byte load per thread (= row) and 16-column chunk, which is latency-bound when the rows come from DRAM.
 * A hint to the memory system: results are unchanged by construction.  `on` is a mask: bit 0 = h of the GEGLU data gradient,
 * bit 1 = residuals; 0 off (default, or NK_GEMM_EPI_PREFETCH); 0..3 sets, anything else queries; returns the previous mask.  Reference sites of the fused adds: modules/attention.py:497-511
 * (x + attn(...), x + ff(...)), modules/diffusion/openaimodel.py:337-342 (skip_connection(x) + h). */
int nk_gemm_set_epi_prefetch(int on);

/* y[M,N] = x[M,K] @ w[N,K]^T (+ bias[N]) (+ residual[M,N]);  y bf16 (out_f32 = 0) or fp32.
 * Replaces nn.Linear forward: modules/attention.py:283-290 (to_q/k/v/to_out), :53,:67-71 (GEGLU /
 * FeedForward), :618,:639 (proj_in/out); modules/diffusion/openaimodel.py:586-590,273-279. */
int nk_linear_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                  const void* residual, int64_t ldr, void* y, int64_t ldy, int out_f32, int M, int N,
                  int K, nk_stream_t stream);
/* dx[M,K] = dy[M,N] @ w[N,K] (+ residual[M,K])  — autograd of nn.Linear w.r.t. its input. */
int nk_linear_dgrad(const void* dy, int64_t lddy, const void* w, int64_t ldw, const void* residual,
                    int64_t ldr, void* dx, int64_t lddx, int M, int N, int K, nk_stream_t stream);
/* dw[N,K] (fp32) (+)= dy[M,N]^T @ x[M,K]        — autograd of nn.Linear w.r.t. its weight.
 * accumulate = 0 overwrites (dw need not be initialised), 1 adds with fp32 red.global. */
int nk_linear_wgrad(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int64_t lddw,
                    int accumulate, int M, int N, int K, nk_stream_t stream);

/* GEGLU feed-forward input projection, gate fused into the GEMM epilogue — replaces `GEGLU.forward`
 * (/root/reference/src/neurosis/modules/attention.py:50-57: `x, gate = self.proj(x).chunk(2, dim=-1); x * F.gelu(gate)`).
 *   h[M, 2D] = x[M,K] @ w[2D,K]^T + bias[2D]     (bf16; optional — null skips the store; saved for the backward)
 *   out[M, D] = h[:, :D] * gelu_erf(h[:, D:])     (bf16)
 * D % 16 == 0; ldh, ldo % 8 == 0. */
int nk_linear_geglu_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* h, int64_t ldh,
                        void* out, int64_t ldo, int M, int D, int K, nk_stream_t stream);
/* Data gradient of the projection that FOLLOWS a GEGLU (`FeedForward.net[2]`, attention.py:60-74) with the GEGLU
 * backward fused into its epilogue: d_out = dy[M,N] @ w[N,D] is never written,
 *   dh[M, :D] = d_out * gelu_erf(h[:, D:])      dh[M, D:] = d_out * h[:, :D] * gelu_erf'(h[:, D:]). */
int nk_linear_dgrad_geglu(const void* dy, int64_t lddy, const void* w, int64_t ldw, const void* h, int64_t ldh, void* dh,
                          int64_t lddh, int M, int N, int D, nk_stream_t stream);

/* 3x3 (pad 1) or 1x1 (pad 0) stride-1 convolution on NHWC bf16 activations as implicit GEMM.
 *   y[n,h,w,co] = sum_{tap,ci} x[n,h+dy,w+dx,ci] * wp[co, tap*Cin+ci] + bias[co] + bias_img[n,co]
 *                 + residual[n,h,w,co]
 * wp is the packed bf16 weight [Cout, ksize*ksize*Cin] produced by nk_conv_pack_weights.
 * x_pix_stride / y_pix_stride are the element strides between consecutive pixels (>= channels),
 * so channel slices of a wider NHWC buffer can be read / written in place.
 * Replaces nn.Conv2d forward: modules/diffusion/openaimodel.py:247-251,281-294,301,623,800,124;
 * modules/diffusion/model.py:100-111 (VAE ResnetBlock convs). */
int nk_conv2d_fwd(const void* x, int64_t x_pix_stride, const void* wp, const float* bias,
                  const float* bias_img, const void* residual, int64_t r_pix_stride, void* y,
                  int64_t y_pix_stride, int nimg, int H, int W, int Cin, int Cout, int ksize,
                  nk_stream_t stream);
/* 3x3 stride-2 convolution forward as implicit GEMM: output pixel (h, w) reads input (2h + ky - pad_t, 2w + kx - pad_l),
 * out-of-image taps are zero (TMA boxes traversed with element stride 2, zero-filled outside the tensor).  pad 1/1 =
 * openaimodel.Downsample (modules/diffusion/openaimodel.py:146-189); pad 0/0 with Ho = (H + 1 - 3) / 2 + 1 =
 * model.Downsample's ConstantPad2d((0,1,0,1)) + conv k3 s2 p0 (modules/diffusion/model.py:65-82).  wp as for
 * nk_conv2d_fwd; Cin is the physical (64-padded) channel count of x. */
int nk_conv2d_stride2_fwd(const void* x, int64_t x_pix_stride, const void* wp, const float* bias, void* y,
                          int64_t y_pix_stride, int nimg, int H, int W, int Cin, int Cout, int ksize, int pad_t,
                          int pad_l, int Ho, int Wo, nk_stream_t stream);
/* dw_packed[Cout, taps, Cin] (fp32) += sum_pixels dy[p,co] * x[p+tap,ci]   (always accumulates;
 * zero the buffer first).  Autograd of nn.Conv2d w.r.t. its weight. */
int nk_conv2d_wgrad(const void* dy, int64_t dy_pix_stride, const void* x, int64_t x_pix_stride,
                    float* dw_packed, int nimg, int H, int W, int Cin, int Cout, int ksize,
                    nk_stream_t stream);

/* Fused flash-style attention forward (tcgen05/TMEM), no mask/dropout, head_dim D = 64 or a multiple of 64 up to 512.
 * q,k,v: bf16 [B, N, H, D] addressed by (batch_stride, row_stride, head offset h*D); they may be
 * column slices of a wider projection output.  o: bf16 [B, Nq, H, D]; lse: fp32 [B, H, Nq] or NULL.
 * D = 64: one CTA per 128 query rows and head, two CTAs per SM.  D > 64 (the single 512-wide head of the VAE
 * mid-block attention, AttnBlock / MemoryEfficientAttnBlock, modules/diffusion/model.py:144-236):
 * scores are accumulated in TMEM over the 64-wide chunks of D and each CTA produces one 64-wide chunk of o.
 * Replaces F.scaled_dot_product_attention / xformers memory_efficient_attention at
 * modules/attention.py:346-352,410-412. */
int nk_attention_fwd(const void* q, int64_t q_row_stride, int64_t q_batch_stride, const void* k,
                     int64_t k_row_stride, int64_t k_batch_stride, const void* v, int64_t v_row_stride,
                     int64_t v_batch_stride, void* o, int64_t o_row_stride, int64_t o_batch_stride,
                     float* lse, int B, int H, int Nq, int Nk, int head_dim, float scale, nk_stream_t stream);
/* Fused attention backward (head_dim 64): dq_acc fp32 [B,Nq,H,64] must be zero-initialised (K/V tiles accumulate
 * into it with TMA reduce-adds); dk, dv bf16 [B,Nk,H,64] with element strides dkv_row_stride / dkv_batch_stride
 * (0 = contiguous; non-zero lets them be column slices of one fused q|k|v gradient buffer); lse from
 * nk_attention_fwd, delta from nk_attn_delta.  Autograd of the attention call sites
 * modules/attention.py:346-352,410-412. */
int nk_attention_bwd(const void* q, int64_t q_row_stride, int64_t q_batch_stride, const void* k, int64_t k_row_stride,
                     int64_t k_batch_stride, const void* v, int64_t v_row_stride, int64_t v_batch_stride,
                     const void* dO, int64_t do_row_stride, int64_t do_batch_stride, const float* lse,
                     const float* delta, float* dq_acc, void* dk, void* dv, int64_t dkv_row_stride,
                     int64_t dkv_batch_stride, int B, int H, int Nq, int Nk, int head_dim, float scale,
                     nk_stream_t stream);
/* delta[b,h,q] = sum_d dO*O on [B,N,H,D] tensors (attention backward). */
int nk_attn_delta(const void* dO, const void* O, float* delta, int B, int N, int H, int D, nk_stream_t stream);
/* P = softmax(scale*S) row-wise (S fp32 [rows, lds], P bf16), lse optional — materialised attention path
 * for head dims the fused kernel does not cover (SD1.5 d=40/80/160, VAE d=512: model.py:155-166). */
int nk_softmax_rows(const float* S, int64_t lds, void* P, int64_t ldp, float* lse, int64_t rows, int N, float scale,
                    nk_stream_t stream);

/* ---- normalisation (HBM-bound) --------------------------------------------------------------- */
/* GroupNorm (+ optional fused SiLU) on NHWC bf16; mean/rstd fp32 [nimg*G] are outputs of fwd and inputs of bwd.
 * Replaces nn.GroupNorm(32,C)+nn.SiLU: modules/diffusion/openaimodel.py:247-249,281-283,798 (eps 1e-5);
 * modules/attention.py:612 and modules/layers.py:5-7 (eps 1e-6). */
int64_t nk_groupnorm_workspace_bytes(int nimg, int HW, int C, int G);
int nk_groupnorm_fwd(const void* x, int64_t x_pix_stride, const float* gamma, const float* beta, void* y,
                     int64_t y_pix_stride, float* mean, float* rstd, void* workspace, int64_t workspace_bytes,
                     int nimg, int HW, int C, int G, float eps, int silu, nk_stream_t stream);
/* dgamma/dbeta (fp32 [C]) are accumulated into (+=); pass NULL to skip. */
int nk_groupnorm_bwd(const void* dy, int64_t dy_pix_stride, const void* x, int64_t x_pix_stride,
                     const float* gamma, const float* beta, const float* mean, const float* rstd, void* dx,
                     int64_t dx_pix_stride, float* dgamma, float* dbeta, void* workspace, int64_t workspace_bytes,
                     int nimg, int HW, int C, int G, int silu, nk_stream_t stream);
/* Kernel forms written after the last GPU run, a bit mask:
 *   bit 0  LayerNorm forward as column-owner blocks (gamma / beta in registers, shifted single-pass variance);
 *   bit 2  LayerNorm backward as column-owner blocks (dgamma / dbeta sums in registers, ONE pass over x and dy);
 *   bit 1  the second pass of GroupNorm forward / backward walks its (image, chunk) grid backwards, starting on the part of
 *          x (and dy) the statistics pass read last and that is still in L2 (openaimodel.py:247-249).
 * 0 = the measured forms (library default, or NK_NORM_VARIANT).  Same formulas, different reduction / dispatch order: outputs
 * agree to bf16 rounding, statistics to fp32 rounding.  mask outside 0..255 only queries; returns the previous mask.
 * neurosis_b200.tune sets bits after an on-device comparison.  Reference counterpart of the kernels: nn.LayerNorm,
 * modules/attention.py:468-470. */
int nk_norm_set_variant(int mask);

/* LayerNorm over the last dim of [rows, C] bf16.  Replaces nn.LayerNorm: modules/attention.py:468-470. */
int nk_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy,
                     float* mean, float* rstd, int rows, int C, float eps, nk_stream_t stream);
/* dgamma/dbeta fp32 [C] accumulated with atomics (+=). */
/* dres (nullable, bf16 [rows, lddres]): gradient that reaches x through the residual branch around the norm
 * (x -> LN -> f(.) + x, modules/attention.py:497-511); it is added into dx, so autograd needs no separate add. */
int nk_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* gamma,
                     const float* mean, const float* rstd, const void* dres, int64_t lddres, void* dx, int64_t lddx,
                     float* dgamma, float* dbeta, int rows, int C, nk_stream_t stream);

/* ---- elementwise / layout (HBM-bound) --------------------------------------------------------- */
/* GEGLU: h = [value | gate] (bf16 [M, 2D]); out = value * gelu_erf(gate).  modules/attention.py:50-57 */
int nk_geglu_fwd(const void* h, int64_t ldh, void* out, int64_t ldo, int64_t M, int D, nk_stream_t stream);
int nk_geglu_bwd(const void* h, int64_t ldh, const void* dout, int64_t ldo, void* dh, int64_t lddh, int64_t M, int D,
                 nk_stream_t stream);
int nk_silu_fwd(const void* x, void* y, int64_t n, nk_stream_t stream);
int nk_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, nk_stream_t stream);
int nk_add(const void* a, const void* b, void* y, int64_t n, nk_stream_t stream);
int nk_cast_f32_bf16(const float* x, void* y, int64_t n, nk_stream_t stream);
/* the per-step fp32 -> bf16 refresh of every matmul weight in ONE launch (the reference gets the same effect from
 * torch.autocast's per-call weight casts, models/diffusion.py:232 / trainer precision "bf16-mixed").  spans_dev: device
 * array of n_spans records {const float* src; bf16* dst; int64 n} (24 bytes each); one thread block per span. */
int nk_cast_f32_bf16_multi(const void* spans_dev, int n_spans, nk_stream_t stream);
int nk_cast_bf16_f32(const void* x, float* y, int64_t n, int accumulate, nk_stream_t stream);
/* row-strided fp32 -> bf16: y[r, 0:cols] = x[r, 0:cols] with leading dimensions ldx / ldy (elements; cols, ldx, ldy
 * multiples of 8): the fp32 dQ accumulator of the attention backward into its slice of a fused q|k|v gradient. */
int nk_cast_f32_bf16_rows(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int cols,
                          nk_stream_t stream);
/* copy C channels of every pixel between NHWC buffers with different pixel strides (skip concat / split:
 * modules/diffusion/openaimodel.py:836) */
int nk_copy_channels(const void* src, int64_t src_stride, void* dst, int64_t dst_stride, int64_t npix, int C,
                     nk_stream_t stream);
/* nearest-neighbour 2x upsampling, NHWC (modules/diffusion/openaimodel.py:140) and its adjoint */
int nk_upsample2x_fwd(const void* x, void* y, int nimg, int H, int W, int C, nk_stream_t stream);
int nk_upsample2x_bwd(const void* dy, void* dx, int nimg, int H, int W, int C, nk_stream_t stream);
/* NCHW (fp32|bf16) -> NHWC bf16 with zero channel padding to Cpad and optional per-image scale; and back */
int nk_nchw_to_nhwc(const void* src, int src_is_f32, void* dst, const float* scale, int nimg, int C, int HW, int Cpad,
                    nk_stream_t stream);
int nk_nhwc_to_nchw(const void* src, int64_t src_stride, void* dst, int dst_is_f32, int nimg, int C, int HW,
                    nk_stream_t stream);
/* 3x3 (pad 1) patches of a thin NCHW fp32 image (C = 1, 3 or 4): col[p, tap*C + c] bf16 [nimg*H*W, 64], 9*C used, rest 0.
 * Feeds the first VAE convolution Encoder.conv_in (3 -> 128, modules/diffusion/model.py:515-517) as a K = 64 GEMM. */
int nk_image_patches3x3(const float* x, void* col, int nimg, int C, int H, int W, nk_stream_t stream);
/* explicit im2col / col2im for the stride-2 convolutions (openaimodel.py:183-190; VAE model.py:65-82 with
 * pad (0,1,0,1) expressed as pad_t = pad_l = 0) */
int nk_im2col(const void* x, int64_t x_stride, void* col, int nimg, int H, int W, int C, int ks, int stride, int pad_t,
              int pad_l, int Ho, int Wo, nk_stream_t stream);
int nk_col2im(const void* dcol, void* dx, int nimg, int H, int W, int C, int ks, int stride, int pad_t, int pad_l,
              int Ho, int Wo, nk_stream_t stream);
/* out[g, c] (=|+=) sum over the rows of group g of x[rows, C] (bias / timestep-embedding gradients) */
int nk_colsum(const void* x, int64_t ldx, float* out, int groups, int rows_per_group, int C, int accumulate,
              nk_stream_t stream);
/* sinusoidal embedding cos|sin (modules/diffusion/util.py:152-177); t fp32 [B], out bf16 [B, dim] */
int nk_timestep_embedding(const float* t, void* out, int B, int dim, float max_period, nk_stream_t stream);
/* fp32 OIHW conv weight -> packed bf16 forward [CoP, taps*CiP] and data-gradient [CiP, taps*CoP] operands */
int nk_conv_pack_weights(const float* w, void* wp_fwd, void* wp_dgrad, int Co, int Ci, int ks, int CoP, int CiP,
                         nk_stream_t stream);
/* packed fp32 weight gradient [Co, taps, CiP] -> OIHW fp32 (=|+=) */
int nk_conv_unpack_wgrad(const float* dw_packed, float* dw, int Co, int Ci, int ks, int CiP, int accumulate,
                         nk_stream_t stream);

/* ---- diffusion objective (modules/diffusion/loss.py:117-155, denoiser.py:40-57) --------------- */
int nk_noise_mix(const float* x, const float* noise, const float* sigma, float* z, int B, int64_t per_sample,
                 int rectified_flow, nk_stream_t stream);
/* y[b,:] = s[b] * x[b,:]  (c_in scaling of the network input, denoiser.py:49) */
int nk_scale_per_sample(const float* x, const float* s, float* y, int B, int64_t per_sample, nk_stream_t stream);
/* out[b,:] = a[b]*x[b,:] + c[b]*y[b,:] (y may be NULL): D = F*c_out + z_t*c_skip, denoiser.py:53 */
int nk_lincomb_per_sample(const float* x, const float* a, const float* y, const float* c, float* out, int B,
                          int64_t per_sample, nk_stream_t stream);
/* loss[b] = w[b] * mean((D[b]-T[b])^2) and its gradient w.r.t. D */
int nk_weighted_mse_fwd(const float* D, const float* T, const float* w, float* loss, int B, int64_t per_sample,
                        nk_stream_t stream);
int nk_weighted_mse_bwd(const float* D, const float* T, const float* w, const float* dloss, float* dD, int B,
                        int64_t per_sample, nk_stream_t stream);
/* the L1 form (BatchL1Loss, modules/losses/functions.py:65-78; StandardDiffusionLoss loss_type "l1"):
 * loss[b] = w[b] * mean(|D[b]-T[b]|), gradient w[b]/n * sign(D - T) */
int nk_weighted_l1_fwd(const float* D, const float* T, const float* w, float* loss, int B, int64_t per_sample,
                       nk_stream_t stream);
int nk_weighted_l1_bwd(const float* D, const float* T, const float* w, const float* dloss, float* dD, int B,
                       int64_t per_sample, nk_stream_t stream);

/* ---- VAE posterior (section 8(f) row 1: VAE training step) ------------------------------------
 * DiagonalGaussianDistribution (modules/distributions.py:29-51) as used by DiagonalGaussianRegularizer
 * (modules/regularizers.py:31-41): moments (B, 2C, H, W) fp32 NCHW, half = C*H*W.  logvar clamped to [-30, 20];
 * z = mean + exp(logvar/2) * eps (eps NULL: mode); kl[b] = 0.5 * sum(mean^2 + var - 1 - logvar) (kl may be NULL). */
int nk_diag_gaussian_fwd(const float* moments, const float* eps, float* z, float* kl, int B, int64_t half,
                         nk_stream_t stream);
/* gradient w.r.t. the moments from dz (B, C, H, W; may be NULL) and dkl (B; may be NULL) */
int nk_diag_gaussian_bwd(const float* moments, const float* eps, const float* dz, const float* dkl, float* dmoments,
                         int B, int64_t half, nk_stream_t stream);

/* ---- optimizer side --------------------------------------
 * One Adafactor step (reference optimizers/adafactor.py:166-246) over ALL parameters in five stream-ordered launches.
 * tensors_dev: device array of n_tensors records (88 bytes each, 8-byte aligned):
 *   { float* p; const float* g; float* vr; float* vc; float* exp_avg; bf16* mirror; float* scratch; int64 n;
 *     int32 kind, Bt, R, C, group, owner; }
 *   kind 0: 1-D parameter, vr = full second moment [n];  kind 1: Bt matrices of R x C <= 64 elements (conv kernels),
 *   vr [Bt*R], vc [Bt*C];  kind 2: one large R x C matrix, vr [R], vc [C], scratch [R + C] floats (zero on entry, left
 *   zero);  n = elements of the whole parameter; owner = index of the parameter's first record (a [Bt, R, C]
 *   parameter with large R*C is Bt kind-2 records sharing their owner's sums);  exp_avg (first moment) and mirror (bf16 copy of p, refreshed in the same pass) may be NULL.
 * blk_start_dev: int32 [n_tensors] first thread block of every tensor (kind 0: ceil(n/4096) blocks, kind 1:
 *   ceil(Bt/1024), kind 2: ceil(R/64)*ceil(C/128)); n_blocks = their total.
 * hyper_dev: float [groups][8] = {beta2t, rel_step, eps1, eps2, clip_threshold, weight_decay, beta1, scale_parameter}
 *   written by the host before every step (beta2t = 1 - step^decay_rate, rel_step = min(1e-6*step | 1e-2, step^-1/2)
 *   or the external lr).  scal_dev: float [n_tensors][4] workspace.  rms_out: float [n_tensors] = RMS(p) before the
 *   update (state["RMS"] of the reference), may be NULL. */
int nk_adafactor_step(const void* tensors_dev, const int32_t* blk_start_dev, int n_tensors, int n_blocks,
                      const float* hyper_dev, float* scal_dev, float* rms_out, nk_stream_t stream);
/* Graph-replayable form of the per-step scalars: advances the device-resident step count of every parameter group
 * (int64 [groups]) and derives hyper_dev from it and the constants consts_dev = float [groups][8] = {decay_rate, lr,
 * eps1, eps2, clip_threshold, weight_decay, beta1, flags (1 relative_step | 2 warmup_init | 4 scale_parameter)}
 * (adafactor.py:131-134, 216).  Launch before nk_adafactor_step inside the captured step. */
int nk_adafactor_hyper(const float* consts_dev, int64_t* step_dev, float* hyper_dev, int n_groups, nk_stream_t stream);
/* LitEma decay warm-up from the device-resident counter (ema.py:43-46): increments *num_updates_dev when it is >= 0 and
 * writes one_minus_decay = 1 - min(decay, (1 + n) / (10 + n)) for nk_ema_update_multi. */
int nk_ema_decay(float decay, int32_t* num_updates_dev, float* one_minus_decay_dev, nk_stream_t stream);
/* EMA shadow update of LitEma.forward (modules/ema.py:40-59): shadow -= (1 - decay) * (shadow - p) for every span
 * {float* shadow; const float* p; int64 n} (24 bytes) of the device table, one thread block per span; the factor is
 * read from device memory so a captured graph can be replayed with a changing decay. */
int nk_ema_update_multi(const void* spans_dev, int n_spans, const float* one_minus_decay_dev, nk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NK_B200_H_ */
