// Internal description of one tcgen05 GEMM launch (shared by the linear, conv and attention
// entry points).  See gemm_tc.cu for the kernel.
#pragma once
#include "common.cuh"

namespace nk {

enum GemmEpilogue : int {
    EPI_LINEAR = 0,  // C = alpha*acc (+ bias[n]) (+ bias_img[img,n]) (+ residual[m,n])
    EPI_EXP2 = 1,    // C = exp2(alpha*acc - rowvec[m])                 (attention probabilities)
    EPI_DSOFTMAX = 2,  // C = aux[m,n] * (acc - rowvec[m]) * alpha       (attention dS)
    // GEGLU fused into the feed-forward GEMMs (reference modules/attention.py:50-57: x, gate = proj(x).chunk(2); x * gelu(gate)).
    // FWD: B = W [2D, K]; an N tile holds BN/2 value columns and the BN/2 gate columns of the same features, so one thread
    //      sees both: C (optional) = h [M, 2D] = acc + bias (saved for the backward), C2 [M, D] = value * gelu(gate).
    // BWD: the data-gradient GEMM of the projection that FOLLOWS the GEGLU: acc = d(out) [M, D]; with aux = h [M, 2D]
    //      C [M, 2D] = (acc * gelu(gate) | acc * value * gelu'(gate)) — d(out) itself is never written.
    EPI_GEGLU_FWD = 3,
    EPI_GEGLU_BWD = 4,
};

enum GemmOut : int {
    OUT_BF16 = 0,
    OUT_F32 = 1,
    OUT_F32_ATOMIC = 2,  // red.global.add.f32 into C (split-K / gradient accumulation)
};

// One operand.  Matrix form: `rows` x `inner` with `inner` contiguous, plus two batch dims.
// Image form (conv=1): NHWC tensor (nimg, H, W, inner) with pixel stride `row_stride`.
struct GemmOperand {
    const bf16* ptr;
    int mn_major;  // 0: `inner` is the reduction dim K;  1: `inner` is the M (or N) dim
    int conv;      // 1: addressed through pixel tiles (4D coordinates c, w, h, img)
    long long inner, rows, row_stride;      // elements
    long long nb2, b2_stride, nb1, b1_stride;  // batch dims (matrix form)
    int H, W, nimg;                         // image form
};

struct GemmProblem {
    GemmOperand A, B;
    int M, N, K;          // per batch entry
    int nb2, nb1;         // batch extents of the launch (C is indexed by them as well)
    // conv forward / dgrad (A.conv=1, A K-major): K = taps*Cin, k-iteration -> (tap, 64-channel block)
    int ksize, pad;       // 3,1 or 1,0
    // strided conv forward (conv_stride = 2): output pixel (h, w) reads input (2h + ky - pad_t, 2w + kx - pad_l); the
    // A tiles are TMA boxes traversed with element stride 2.  0 / 1 = ordinary stride-1 convolution (pad both sides).
    int conv_stride, pad_t, pad_l, out_H, out_W;
    // conv wgrad (A.conv = B.conv = 1, both MN-major): N = taps*Cin (every 64-channel atom has its own tap shift), K = all pixels
    int wgrad;
    // output
    void* C;
    long long ldc, c_b2_stride, c_b1_stride;  // elements
    int out;              // GemmOut
    int epi;              // GemmEpilogue
    float alpha;
    const float* bias;        // [N] or null
    const float* bias_img;    // [nimg, N] or null (row m belongs to image m / rows_per_img)
    int rows_per_img;
    const bf16* residual;     // [M, ldr] or null (batch strides = C's)
    long long ldr;
    const float* rowvec;      // [M] per batch entry (stride M), EPI_EXP2 / EPI_DSOFTMAX
    const bf16* aux;          // [M, N] like C, EPI_DSOFTMAX; h [M, 2D] with row stride ldr for EPI_GEGLU_BWD
    int geglu_d;              // D of the GEGLU epilogues (N = D for both: the tile scheduler runs over D)
    void* C2;                 // EPI_GEGLU_FWD: gated output [M, D]
    long long ldc2;
    int force_bn;             // 0 = heuristic
    int force_splits;         // 0 = heuristic (only with OUT_F32_ATOMIC)
    int force_cta_group;      // 0 = heuristic, 1 = single CTA tiles, 2 = CTA pairs (cta_group::2)
    int force_dual;           // 0 = global mode (gemm_set_dual / NK_GEMM_DUAL), -1 = single row tiles, 1 = cost model, 2 = pair wherever legal
};

int launch_gemm(const GemmProblem& p, cudaStream_t stream);
// row-tile pairing mode of launch_gemm (see gemm_tc.cu): 0 off, 1 cost model, 2 wherever legal; returns the previous mode
int gemm_set_dual(int mode);
// mode 1 only: smallest number of 64-deep k-iterations a launch must have to be paired (< 0 queries); returns the previous value
int gemm_set_dual_min_k(int k_iters);
// mode 1 only: launch classes that may pair (bit 0 K-major-A matrix GEMM, bit 1 MN-major-A matrix GEMM, bit 2 convolution); 0..7 sets, else queries
int gemm_set_dual_classes(int mask);
// L2 prefetch of the epilogue's side input at tile start: mask (bit 0 GEGLU h, bit 1 residual), 0..3 sets, anything else queries; returns the previous mask
int gemm_set_epi_prefetch(int on);
// paired launches: k-iterations by which the second row tile trails the first (0..7, clamped to stages - 1; < 0 queries)
int gemm_set_dual_skew(int k_iters);

}  // namespace nk
