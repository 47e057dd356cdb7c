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
    int32_t wgrad;      /* conv weight gradient: A.conv = B.conv = 1, b2 = filter tap */
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
} nk_gemm_desc;

int nk_gemm_ex(const nk_gemm_desc* d, nk_stream_t stream);

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
/* dw_packed[Cout, taps, Cin] (fp32) += sum_pixels dy[p,co] * x[p+tap,ci]   (always accumulates;
 * zero the buffer first).  Autograd of nn.Conv2d w.r.t. its weight. */
int nk_conv2d_wgrad(const void* dy, int64_t dy_pix_stride, const void* x, int64_t x_pix_stride,
                    float* dw_packed, int nimg, int H, int W, int Cin, int Cout, int ksize,
                    nk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NK_B200_H_ */
