// C-ABI entry points of the tensor-core GEMM family (declared in include/nk_b200.h).
#include "../../include/nk_b200.h"
#include "gemm_tc.cuh"

using namespace nk;

static GemmOperand to_operand(const nk_operand& o) {
    GemmOperand r;
    r.ptr = static_cast<const bf16*>(o.ptr);
    r.mn_major = o.mn_major;
    r.conv = o.conv;
    r.inner = o.inner;
    r.rows = o.rows;
    r.row_stride = o.row_stride;
    r.nb2 = o.nb2;
    r.b2_stride = o.b2_stride;
    r.nb1 = o.nb1;
    r.b1_stride = o.b1_stride;
    r.H = o.H;
    r.W = o.W;
    r.nimg = o.nimg;
    return r;
}

static GemmOperand matrix(const void* p, int mn_major, long long rows, long long inner, long long ld) {
    GemmOperand r;
    memset(&r, 0, sizeof(r));
    r.ptr = static_cast<const bf16*>(p);
    r.mn_major = mn_major;
    r.inner = inner;
    r.rows = rows;
    r.row_stride = ld;
    r.nb2 = r.nb1 = 1;
    return r;
}

static GemmOperand image(const void* p, int mn_major, int nimg, int H, int W, long long C, long long pix_stride) {
    GemmOperand r;
    memset(&r, 0, sizeof(r));
    r.ptr = static_cast<const bf16*>(p);
    r.mn_major = mn_major;
    r.conv = 1;
    r.inner = C;
    r.row_stride = pix_stride;
    r.H = H;
    r.W = W;
    r.nimg = nimg;
    r.nb2 = r.nb1 = 1;
    return r;
}

static GemmProblem blank_problem() {
    GemmProblem p;
    memset(&p, 0, sizeof(p));
    p.nb2 = p.nb1 = 1;
    p.ksize = 1;
    p.alpha = 1.f;
    p.rows_per_img = 1;
    return p;
}

extern "C" {

int nk_version(void) { return 100; }
const char* nk_last_error(void) { return nk::last_error(); }
int nk_sm_count(void) { return nk::device_sm_count(); }
int nk_gemm_set_dual(int mode) { return nk::gemm_set_dual(mode); }
int nk_gemm_set_dual_min_k(int k_iters) { return nk::gemm_set_dual_min_k(k_iters); }
int nk_gemm_set_dual_classes(int mask) { return nk::gemm_set_dual_classes(mask); }
int nk_gemm_set_dual_skew(int k_iters) { return nk::gemm_set_dual_skew(k_iters); }
int nk_gemm_set_epi_prefetch(int on) { return nk::gemm_set_epi_prefetch(on); }

int nk_gemm_ex(const nk_gemm_desc* d, nk_stream_t stream) {
    NK_REQUIRE(d != nullptr, NK_ERR_SHAPE, "null descriptor");
    GemmProblem p = blank_problem();
    p.A = to_operand(d->A);
    p.B = to_operand(d->B);
    p.M = d->M;
    p.N = d->N;
    p.K = d->K;
    p.nb2 = d->nb2 > 0 ? d->nb2 : 1;
    p.nb1 = d->nb1 > 0 ? d->nb1 : 1;
    p.ksize = d->ksize > 0 ? d->ksize : 1;
    p.pad = d->pad;
    p.wgrad = d->wgrad;
    p.C = d->C;
    p.ldc = d->ldc;
    p.c_b2_stride = d->c_b2_stride;
    p.c_b1_stride = d->c_b1_stride;
    p.out = d->out;
    p.epi = d->epi;
    p.alpha = d->alpha;
    p.rows_per_img = d->rows_per_img;
    p.bias = d->bias;
    p.bias_img = d->bias_img;
    p.residual = static_cast<const bf16*>(d->residual);
    p.ldr = d->ldr;
    p.rowvec = d->rowvec;
    p.aux = static_cast<const bf16*>(d->aux);
    p.force_bn = d->force_bn;
    p.force_splits = d->force_splits;
    p.force_cta_group = d->force_cta_group;
    p.force_dual = d->force_dual;
    return launch_gemm(p, ::nk::enter(stream));
}

int nk_linear_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                  const void* residual, int64_t ldr, void* y, int64_t ldy, int out_f32, int M, int N,
                  int K, nk_stream_t stream) {
    GemmProblem p = blank_problem();
    p.A = matrix(x, 0, M, K, ldx);
    p.B = matrix(w, 0, N, K, ldw);
    p.M = M;
    p.N = N;
    p.K = K;
    p.C = y;
    p.ldc = ldy;
    p.out = out_f32 ? OUT_F32 : OUT_BF16;
    p.epi = EPI_LINEAR;
    p.bias = bias;
    p.residual = static_cast<const bf16*>(residual);
    p.ldr = ldr;
    return launch_gemm(p, ::nk::enter(stream));
}

int nk_linear_dgrad(const void* dy, int64_t lddy, const void* w, int64_t ldw, const void* residual,
                    int64_t ldr, void* dx, int64_t lddx, int M, int N, int K, nk_stream_t stream) {
    // dx[m,k] = sum_n dy[m,n] w[n,k]: reduction over n; B[k, n] = w[n,k] is MN-major (k contiguous)
    GemmProblem p = blank_problem();
    p.A = matrix(dy, 0, M, N, lddy);
    p.B = matrix(w, 1, N, K, ldw);
    p.M = M;
    p.N = K;
    p.K = N;
    p.C = dx;
    p.ldc = lddx;
    p.out = OUT_BF16;
    p.epi = EPI_LINEAR;
    p.residual = static_cast<const bf16*>(residual);
    p.ldr = ldr;
    return launch_gemm(p, ::nk::enter(stream));
}

int nk_linear_geglu_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* h, int64_t ldh,
                        void* out, int64_t ldo, int M, int D, int K, nk_stream_t stream) {
    GemmProblem p = blank_problem();
    p.A = matrix(x, 0, M, K, ldx);
    p.B = matrix(w, 0, 2LL * D, K, ldw);
    p.M = M;
    p.N = D;
    p.K = K;
    p.C = h;
    p.ldc = ldh > 0 ? ldh : 2LL * D;
    p.C2 = out;
    p.ldc2 = ldo;
    p.geglu_d = D;
    p.out = OUT_BF16;
    p.epi = EPI_GEGLU_FWD;
    p.bias = bias;
    return launch_gemm(p, ::nk::enter(stream));
}

int nk_linear_dgrad_geglu(const void* dy, int64_t lddy, const void* w, int64_t ldw, const void* h, int64_t ldh, void* dh,
                          int64_t lddh, int M, int N, int D, nk_stream_t stream) {
    // d(out)[m,d] = sum_n dy[m,n] w[n,d] (w = the [N, D] weight of the projection after the GEGLU), turned into
    // dh [M, 2D] by the epilogue
    GemmProblem p = blank_problem();
    p.A = matrix(dy, 0, M, N, lddy);
    p.B = matrix(w, 1, N, D, ldw);
    p.M = M;
    p.N = D;
    p.K = N;
    p.C = dh;
    p.ldc = lddh;
    p.geglu_d = D;
    p.aux = static_cast<const bf16*>(h);
    p.ldr = ldh;
    p.out = OUT_BF16;
    p.epi = EPI_GEGLU_BWD;
    return launch_gemm(p, ::nk::enter(stream));
}

int nk_linear_wgrad(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int64_t lddw,
                    int accumulate, int M, int N, int K, nk_stream_t stream) {
    // dw[n,k] = sum_m dy[m,n] x[m,k]: reduction over tokens m; both operands MN-major
    GemmProblem p = blank_problem();
    p.A = matrix(dy, 1, M, N, lddy);
    p.B = matrix(x, 1, M, K, ldx);
    p.M = N;
    p.N = K;
    p.K = M;
    p.C = dw;
    p.ldc = lddw;
    p.out = accumulate ? OUT_F32_ATOMIC : OUT_F32;
    p.epi = EPI_LINEAR;
    if (!accumulate) {
        // few output tiles but a long reduction (e.g. 640x640 weights, 32k tokens): zero the output and split K
        // across CTAs with fp32 red.global.add instead of leaving most SMs idle
        const long long tiles = static_cast<long long>((N + 255) / 256) * ((K + 255) / 256);
        if (tiles * 2 <= nk::device_sm_count() / 2 && M >= 2048) {
            NK_CUDA(cudaMemset2DAsync(dw, static_cast<size_t>(lddw) * 4, 0, static_cast<size_t>(K) * 4, N,
                                      ::nk::enter(stream)));
            p.out = OUT_F32_ATOMIC;
        }
    }
    return launch_gemm(p, ::nk::enter(stream));
}

int nk_conv2d_fwd(const void* x, int64_t x_pix_stride, const void* wp, const float* bias,
                  const float* bias_img, const void* residual, int64_t r_pix_stride, void* y,
                  int64_t y_pix_stride, int nimg, int H, int W, int Cin, int Cout, int ksize,
                  nk_stream_t stream) {
    NK_REQUIRE(ksize == 1 || ksize == 3, NK_ERR_UNSUPPORTED, "conv ksize %d", ksize);
    GemmProblem p = blank_problem();
    const int taps = ksize * ksize;
    p.A = image(x, 0, nimg, H, W, Cin, x_pix_stride);
    p.B = matrix(wp, 0, Cout, static_cast<long long>(taps) * Cin, static_cast<long long>(taps) * Cin);
    p.M = nimg * H * W;
    p.N = Cout;
    p.K = taps * Cin;
    p.ksize = ksize;
    p.pad = ksize / 2;
    p.C = y;
    p.ldc = y_pix_stride;
    p.out = OUT_BF16;
    p.epi = EPI_LINEAR;
    p.bias = bias;
    p.bias_img = bias_img;
    p.rows_per_img = H * W;
    p.residual = static_cast<const bf16*>(residual);
    p.ldr = r_pix_stride;
    return launch_gemm(p, ::nk::enter(stream));
}

int nk_conv2d_stride2_fwd(const void* x, int64_t x_pix_stride, const void* wp, const float* bias, void* y,
                          int64_t y_pix_stride, int nimg, int H, int W, int Cin, int Cout, int ksize, int pad_t,
                          int pad_l, int Ho, int Wo, nk_stream_t stream) {
    NK_REQUIRE(ksize == 3, NK_ERR_UNSUPPORTED, "strided conv ksize %d", ksize);
    NK_REQUIRE(Ho > 0 && Wo > 0 && 2 * (Ho - 1) - pad_t < H && 2 * (Wo - 1) - pad_l < W, NK_ERR_SHAPE,
               "strided conv: output %dx%d does not fit input %dx%d", Ho, Wo, H, W);
    GemmProblem p = blank_problem();
    const int taps = ksize * ksize;
    p.A = image(x, 0, nimg, H, W, Cin, x_pix_stride);
    p.B = matrix(wp, 0, Cout, static_cast<long long>(taps) * Cin, static_cast<long long>(taps) * Cin);
    p.M = nimg * Ho * Wo;
    p.N = Cout;
    p.K = taps * Cin;
    p.ksize = ksize;
    p.conv_stride = 2;
    p.pad_t = pad_t;
    p.pad_l = pad_l;
    p.out_H = Ho;
    p.out_W = Wo;
    p.C = y;
    p.ldc = y_pix_stride;
    p.out = OUT_BF16;
    p.epi = EPI_LINEAR;
    p.bias = bias;
    p.rows_per_img = Ho * Wo;
    return launch_gemm(p, ::nk::enter(stream));
}

int nk_conv2d_wgrad(const void* dy, int64_t dy_pix_stride, const void* x, int64_t x_pix_stride,
                    float* dw_packed, int nimg, int H, int W, int Cin, int Cout, int ksize,
                    nk_stream_t stream) {
    NK_REQUIRE(ksize == 1 || ksize == 3, NK_ERR_UNSUPPORTED, "conv ksize %d", ksize);
    GemmProblem p = blank_problem();
    const int taps = ksize * ksize;
    p.A = image(dy, 1, nimg, H, W, Cout, dy_pix_stride);
    p.B = image(x, 1, nimg, H, W, Cin, x_pix_stride);
    p.M = Cout;
    p.N = taps * Cin;  // column index = tap*Cin + ci: the packed gradient is one [Cout, taps*Cin] matrix
    p.K = nimg * H * W;
    p.ksize = ksize;
    p.pad = ksize / 2;
    p.wgrad = 1;
    p.C = dw_packed;
    p.ldc = static_cast<long long>(taps) * Cin;
    p.out = OUT_F32_ATOMIC;
    p.epi = EPI_LINEAR;
    return launch_gemm(p, ::nk::enter(stream));
}

}  // extern "C"
