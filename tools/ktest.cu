// Standalone GPU self-test + micro-benchmark of the tensor-core GEMM family through the C ABI.
// Build: see Makefile target `ktest`.  Usage: ktest <case> | ktest list | ktest all
// Each case compares against an obviously-correct naive CUDA kernel on the same device.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../include/nk_b200.h"

typedef __nv_bfloat16 bf16;

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)

__global__ void fill_bf16(bf16* p, size_t n, uint32_t seed, float scale) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = (uint32_t)i * 2654435761u + seed * 40503u + 12345u;
    x ^= x >> 15;
    x *= 2246822519u;
    x ^= x >> 13;
    x *= 3266489917u;
    x ^= x >> 16;
    float f = ((x & 0xffff) / 65535.0f - 0.5f) * 2.0f * scale;
    p[i] = __float2bfloat16(f);
}
__global__ void fill_f32(float* p, size_t n, uint32_t seed, float scale) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = (uint32_t)i * 2654435761u + seed * 977u + 777u;
    x ^= x >> 15;
    x *= 2246822519u;
    x ^= x >> 13;
    p[i] = ((x & 0xffff) / 65535.0f - 0.5f) * 2.0f * scale;
}

// C[b][m][n] = sum_k A[b](m,k) * B[b](n,k), generic element strides
__global__ void ref_gemm(const bf16* A, long long sam, long long sak, long long sab, const bf16* B,
                         long long sbn, long long sbk, long long sbb, float* C, int M, int N, int K,
                         int nbatch) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int m = blockIdx.y;
    int b = blockIdx.z;
    if (n >= N) return;
    float acc = 0.f;
    const bf16* a = A + b * sab + m * sam;
    const bf16* bb = B + b * sbb + n * sbn;
    for (int k = 0; k < K; ++k) acc += __bfloat162float(a[k * sak]) * __bfloat162float(bb[k * sbk]);
    C[((long long)b * M + m) * N + n] = acc;
}

// y[n,h,w,co] = sum x[n,h+dy,w+dx,ci] * wp[co, tap*Cin+ci]
__global__ void ref_conv(const bf16* x, const bf16* wp, float* y, int nimg, int H, int W, int Cin,
                         int Cout, int ks) {
    int co = blockIdx.y * blockDim.x + threadIdx.x;
    long long pix = blockIdx.x;  // grid.x: up to 2^31-1 pixels
    if (co >= Cout) return;
    int w = pix % W;
    int h = (pix / W) % H;
    int n = pix / ((long long)W * H);
    int pad = ks / 2;
    float acc = 0.f;
    for (int t = 0; t < ks * ks; ++t) {
        int hh = h + t / ks - pad, ww = w + t % ks - pad;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        const bf16* xp = x + (((long long)n * H + hh) * W + ww) * Cin;
        const bf16* wpp = wp + ((long long)co * ks * ks + t) * Cin;
        for (int ci = 0; ci < Cin; ++ci) acc += __bfloat162float(xp[ci]) * __bfloat162float(wpp[ci]);
    }
    y[pix * Cout + co] = acc;
}

// dw[co, tap, ci] = sum_p dy[p,co] * x[p+tap, ci]
__global__ void ref_wgrad(const bf16* dy, const bf16* x, float* dw, int nimg, int H, int W, int Cin,
                          int Cout, int ks) {
    int ci = blockIdx.x * blockDim.x + threadIdx.x;
    int t = blockIdx.y;
    int co = blockIdx.z;
    if (ci >= Cin) return;
    int pad = ks / 2;
    float acc = 0.f;
    for (int n = 0; n < nimg; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                int hh = h + t / ks - pad, ww = w + t % ks - pad;
                if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                acc += __bfloat162float(dy[(((long long)n * H + h) * W + w) * Cout + co]) *
                       __bfloat162float(x[(((long long)n * H + hh) * W + ww) * Cin + ci]);
            }
    dw[((long long)co * ks * ks + t) * Cin + ci] = acc;
}

template <typename T>
T* dalloc(size_t n) {
    T* p;
    CK(cudaMalloc(&p, n * sizeof(T)));
    return p;
}
static void fillb(bf16* p, size_t n, uint32_t seed, float scale = 1.f) {
    fill_bf16<<<(unsigned)((n + 255) / 256), 256>>>(p, n, seed, scale);
}
static void fillf(float* p, size_t n, uint32_t seed, float scale = 1.f) {
    fill_f32<<<(unsigned)((n + 255) / 256), 256>>>(p, n, seed, scale);
}

struct Cmp {
    double max_abs = 0, max_ref = 0;
    long long bad = 0, first_bad = -1;
};

static Cmp compare(const std::vector<float>& got, const std::vector<float>& ref, double atol, double rtol) {
    Cmp c;
    for (size_t i = 0; i < ref.size(); ++i) {
        double d = fabs((double)got[i] - (double)ref[i]);
        if (d > c.max_abs || d != d) c.max_abs = (d != d) ? 1e30 : d;
        if (fabs(ref[i]) > c.max_ref) c.max_ref = fabs(ref[i]);
        if (!(d <= atol + rtol * fabs(ref[i]))) {
            if (c.first_bad < 0) c.first_bad = (long long)i;
            c.bad++;
        }
    }
    return c;
}

static std::vector<float> fetch_bf16(const bf16* d, size_t n) {
    std::vector<bf16> h(n);
    CK(cudaMemcpy(h.data(), d, n * sizeof(bf16), cudaMemcpyDeviceToHost));
    std::vector<float> f(n);
    for (size_t i = 0; i < n; ++i) f[i] = __bfloat162float(h[i]);
    return f;
}
static std::vector<float> fetch_f32(const float* d, size_t n) {
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost));
    return h;
}

static int report(const char* name, const Cmp& c, size_t n, int cols, double ms, double flops) {
    bool ok = c.bad == 0;
    printf("%-40s %s max_abs=%.4g max_ref=%.4g bad=%lld/%zu", name, ok ? "PASS" : "FAIL", c.max_abs,
           c.max_ref, c.bad, n);
    if (!ok) printf(" first_bad=(row %lld, col %lld)", c.first_bad / cols, c.first_bad % cols);
    if (ms > 0) printf(" time=%.3f ms  %.1f TFLOP/s", ms, flops / ms * 1e-9);
    printf("\n");
    fflush(stdout);
    return ok ? 0 : 1;
}

template <typename F>
static double time_ms(F&& f, int iters = 10) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

static void must(int rc, const char* what) {
    if (rc != 0) {
        printf("%s failed rc=%d: %s\n", what, rc, nk_last_error());
        exit(3);
    }
}

// ---------------------------------------------------------------------------------------------
static int test_linear(const char* name, int M, int N, int K, bool bias, bool resid, bool f32out, int bn = 0) {
    bf16* x = dalloc<bf16>((size_t)M * K);
    bf16* w = dalloc<bf16>((size_t)N * K);
    bf16* r = dalloc<bf16>((size_t)M * N);
    float* b = dalloc<float>(N);
    void* y = f32out ? (void*)dalloc<float>((size_t)M * N) : (void*)dalloc<bf16>((size_t)M * N);
    float* ref = dalloc<float>((size_t)M * N);
    fillb(x, (size_t)M * K, 1);
    fillb(w, (size_t)N * K, 2, 0.25f);
    fillb(r, (size_t)M * N, 3);
    fillf(b, N, 4);
    ref_gemm<<<dim3((N + 127) / 128, M, 1), 128>>>(x, K, 1, 0, w, K, 1, 0, ref, M, N, K, 1);
    CK(cudaDeviceSynchronize());
    auto run = [&]() {
        if (bn == 0) {
            must(nk_linear_fwd(x, K, w, K, bias ? b : nullptr, resid ? r : nullptr, N, y, N, f32out, M, N, K, 0), name);
        } else {
            nk_gemm_desc d;
            memset(&d, 0, sizeof(d));
            d.A.ptr = x; d.A.inner = K; d.A.rows = M; d.A.row_stride = K; d.A.nb1 = d.A.nb2 = 1;
            d.B.ptr = w; d.B.inner = K; d.B.rows = N; d.B.row_stride = K; d.B.nb1 = d.B.nb2 = 1;
            d.M = M; d.N = N; d.K = K; d.nb1 = d.nb2 = 1; d.ksize = 1;
            d.C = y; d.ldc = N; d.out = f32out ? NK_OUT_F32 : NK_OUT_BF16; d.alpha = 1.f;
            d.bias = bias ? b : nullptr; d.residual = resid ? r : nullptr; d.ldr = N; d.force_bn = bn;
            must(nk_gemm_ex(&d, 0), name);
        }
    };
    double ms = time_ms(run);
    std::vector<float> got = f32out ? fetch_f32((float*)y, (size_t)M * N) : fetch_bf16((bf16*)y, (size_t)M * N);
    std::vector<float> rf = fetch_f32(ref, (size_t)M * N);
    if (bias || resid) {
        std::vector<float> hb = fetch_f32(b, N), hr = fetch_bf16(r, (size_t)M * N);
        for (size_t i = 0; i < rf.size(); ++i) rf[i] += (bias ? hb[i % N] : 0.f) + (resid ? hr[i] : 0.f);
    }
    Cmp c = compare(got, rf, f32out ? 2e-3 : 3e-2, f32out ? 1e-4 : 1e-2);
    int rc = report(name, c, rf.size(), N, ms, 2.0 * M * N * K);
    cudaFree(x); cudaFree(w); cudaFree(r); cudaFree(b); cudaFree(y); cudaFree(ref);
    return rc;
}

static int test_dgrad(const char* name, int M, int N, int K) {
    bf16* dy = dalloc<bf16>((size_t)M * N);
    bf16* w = dalloc<bf16>((size_t)N * K);
    bf16* dx = dalloc<bf16>((size_t)M * K);
    float* ref = dalloc<float>((size_t)M * K);
    fillb(dy, (size_t)M * N, 5);
    fillb(w, (size_t)N * K, 6, 0.25f);
    // dx[m,k] = sum_n dy[m,n] w[n,k]  -> "A"(m, n) strides (N,1), "B"(k, n) strides (1, K)
    ref_gemm<<<dim3((K + 127) / 128, M, 1), 128>>>(dy, N, 1, 0, w, 1, K, 0, ref, M, K, N, 1);
    CK(cudaDeviceSynchronize());
    double ms = time_ms([&]() { must(nk_linear_dgrad(dy, N, w, K, nullptr, 0, dx, K, M, N, K, 0), name); });
    Cmp c = compare(fetch_bf16(dx, (size_t)M * K), fetch_f32(ref, (size_t)M * K), 3e-2, 1e-2);
    int rc = report(name, c, (size_t)M * K, K, ms, 2.0 * M * N * K);
    cudaFree(dy); cudaFree(w); cudaFree(dx); cudaFree(ref);
    return rc;
}

static int test_wgrad(const char* name, int M, int N, int K, int accumulate) {
    bf16* dy = dalloc<bf16>((size_t)M * N);
    bf16* x = dalloc<bf16>((size_t)M * K);
    float* dw = dalloc<float>((size_t)N * K);
    float* ref = dalloc<float>((size_t)N * K);
    fillb(dy, (size_t)M * N, 7);
    fillb(x, (size_t)M * K, 8);
    // dw[n,k] = sum_m dy[m,n] x[m,k] -> "A"(n, m) strides (1, N), "B"(k, m) strides (1, K)
    ref_gemm<<<dim3((K + 127) / 128, N, 1), 128>>>(dy, 1, N, 0, x, 1, K, 0, ref, N, K, M, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(dw, 0, (size_t)N * K * 4));
    must(nk_linear_wgrad(dy, N, x, K, dw, K, accumulate, M, N, K, 0), name);
    CK(cudaDeviceSynchronize());
    Cmp c = compare(fetch_f32(dw, (size_t)N * K), fetch_f32(ref, (size_t)N * K), 2e-2, 2e-4);
    double ms = time_ms([&]() { must(nk_linear_wgrad(dy, N, x, K, dw, K, accumulate, M, N, K, 0), name); });
    int rc = report(name, c, (size_t)N * K, K, ms, 2.0 * M * N * K);
    cudaFree(dy); cudaFree(x); cudaFree(dw); cudaFree(ref);
    return rc;
}

static int test_conv(const char* name, int nimg, int H, int W, int Cin, int Cout, int ks, bool extras) {
    size_t npix = (size_t)nimg * H * W;
    bf16* x = dalloc<bf16>(npix * Cin);
    bf16* wp = dalloc<bf16>((size_t)Cout * ks * ks * Cin);
    bf16* y = dalloc<bf16>(npix * Cout);
    bf16* r = dalloc<bf16>(npix * Cout);
    float* b = dalloc<float>(Cout);
    float* bi = dalloc<float>((size_t)nimg * Cout);
    float* ref = dalloc<float>(npix * Cout);
    fillb(x, npix * Cin, 9);
    fillb(wp, (size_t)Cout * ks * ks * Cin, 10, 0.1f);
    fillb(r, npix * Cout, 11);
    fillf(b, Cout, 12);
    fillf(bi, (size_t)nimg * Cout, 13);
    ref_conv<<<dim3((unsigned)npix, (Cout + 127) / 128), 128>>>(x, wp, ref, nimg, H, W, Cin, Cout, ks);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    double ms = time_ms([&]() {
        must(nk_conv2d_fwd(x, Cin, wp, extras ? b : nullptr, extras ? bi : nullptr, extras ? r : nullptr, Cout, y,
                           Cout, nimg, H, W, Cin, Cout, ks, 0),
             name);
    });
    std::vector<float> rf = fetch_f32(ref, npix * Cout);
    if (extras) {
        std::vector<float> hb = fetch_f32(b, Cout), hbi = fetch_f32(bi, (size_t)nimg * Cout), hr = fetch_bf16(r, npix * Cout);
        for (size_t i = 0; i < rf.size(); ++i) {
            size_t pix = i / Cout, co = i % Cout, n = pix / ((size_t)H * W);
            rf[i] += hb[co] + hbi[n * Cout + co] + hr[i];
        }
    }
    Cmp c = compare(fetch_bf16(y, npix * Cout), rf, 3e-2, 1e-2);
    int rc = report(name, c, rf.size(), Cout, ms, 2.0 * npix * Cout * ks * ks * Cin);
    cudaFree(x); cudaFree(wp); cudaFree(y); cudaFree(r); cudaFree(b); cudaFree(bi); cudaFree(ref);
    return rc;
}

static int test_conv_wgrad(const char* name, int nimg, int H, int W, int Cin, int Cout, int ks) {
    size_t npix = (size_t)nimg * H * W;
    size_t nw = (size_t)Cout * ks * ks * Cin;
    bf16* x = dalloc<bf16>(npix * Cin);
    bf16* dy = dalloc<bf16>(npix * Cout);
    float* dw = dalloc<float>(nw);
    float* ref = dalloc<float>(nw);
    fillb(x, npix * Cin, 14);
    fillb(dy, npix * Cout, 15, 0.25f);
    ref_wgrad<<<dim3((Cin + 63) / 64, ks * ks, Cout), 64>>>(dy, x, ref, nimg, H, W, Cin, Cout, ks);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(dw, 0, nw * 4));
    must(nk_conv2d_wgrad(dy, Cout, x, Cin, dw, nimg, H, W, Cin, Cout, ks, 0), name);
    CK(cudaDeviceSynchronize());
    Cmp c = compare(fetch_f32(dw, nw), fetch_f32(ref, nw), 3e-2, 3e-4);
    double ms = time_ms([&]() { must(nk_conv2d_wgrad(dy, Cout, x, Cin, dw, nimg, H, W, Cin, Cout, ks, 0), name); });
    int rc = report(name, c, nw, ks * ks * Cin, ms, 2.0 * npix * Cout * ks * ks * Cin);
    cudaFree(x); cudaFree(dy); cudaFree(dw); cudaFree(ref);
    return rc;
}

// batched attention-shaped GEMM: S[b,h] = Q[b,:,h,:] K[b,:,h,:]^T on a [B, N, H, D] layout
static int test_batched_qk(const char* name, int B, int Hh, int Nq, int Nk, int D) {
    size_t nq = (size_t)B * Nq * Hh * D, nk = (size_t)B * Nk * Hh * D, ns = (size_t)B * Hh * Nq * Nk;
    bf16* q = dalloc<bf16>(nq);
    bf16* k = dalloc<bf16>(nk);
    bf16* s = dalloc<bf16>(ns);
    float* ref = dalloc<float>(ns);
    fillb(q, nq, 16);
    fillb(k, nk, 17);
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < Hh; ++h)
            ref_gemm<<<dim3((Nk + 127) / 128, Nq, 1), 128>>>(q + ((size_t)b * Nq * Hh + h) * D, (long long)Hh * D, 1, 0,
                                                             k + ((size_t)b * Nk * Hh + h) * D, (long long)Hh * D, 1, 0,
                                                             ref + ((size_t)b * Hh + h) * Nq * Nk, Nq, Nk, D, 1);
    CK(cudaDeviceSynchronize());
    nk_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.A.ptr = q; d.A.inner = D; d.A.rows = Nq; d.A.row_stride = (int64_t)Hh * D;
    d.A.nb2 = Hh; d.A.b2_stride = D; d.A.nb1 = B; d.A.b1_stride = (int64_t)Nq * Hh * D;
    d.B.ptr = k; d.B.inner = D; d.B.rows = Nk; d.B.row_stride = (int64_t)Hh * D;
    d.B.nb2 = Hh; d.B.b2_stride = D; d.B.nb1 = B; d.B.b1_stride = (int64_t)Nk * Hh * D;
    d.M = Nq; d.N = Nk; d.K = D; d.nb2 = Hh; d.nb1 = B; d.ksize = 1;
    d.C = s; d.ldc = Nk; d.c_b2_stride = (int64_t)Nq * Nk; d.c_b1_stride = (int64_t)Hh * Nq * Nk;
    d.out = NK_OUT_BF16; d.alpha = 1.f;
    double ms = time_ms([&]() { must(nk_gemm_ex(&d, 0), name); });
    Cmp c = compare(fetch_bf16(s, ns), fetch_f32(ref, ns), 3e-2, 1e-2);
    int rc = report(name, c, ns, Nk, ms, 2.0 * B * Hh * Nq * Nk * D);
    cudaFree(q); cudaFree(k); cudaFree(s); cudaFree(ref);
    return rc;
}

struct Case {
    const char* name;
    int (*fn)();
};

static Case cases[] = {
    {"lin_small", []() { return test_linear("lin_small 128x128x64", 128, 128, 64, false, false, true); }},
    {"lin_k256", []() { return test_linear("lin_k256 128x256x256", 128, 256, 256, false, false, true); }},
    {"lin_bn64", []() { return test_linear("lin_bn64 256x128x128 bn=64", 256, 128, 128, false, false, true, 64); }},
    {"lin_bn16", []() { return test_linear("lin_bn16 256x48x128 bn=16", 256, 48, 128, false, false, true, 16); }},
    {"lin_ragged", []() { return test_linear("lin_ragged 300x328x200", 300, 328, 200, true, true, false); }},
    {"lin_multi", []() { return test_linear("lin_multi 1024x1280x320", 1024, 1280, 320, true, true, false); }},
    {"lin_persist", []() { return test_linear("lin_persist 4096x2560x640 (many tiles/CTA)", 4096, 2560, 640, true, false, false); }},
    {"lin_tiny_n", []() { return test_linear("lin_tiny_n 512x4x320", 512, 4, 320, true, false, false); }},
    {"lin_m2", []() { return test_linear("lin_m2 2x1280x320 (embed MLP)", 2, 1280, 320, true, false, true); }},
    {"lin_big", []() { return test_linear("lin_big 8192x1280x1280", 8192, 1280, 1280, true, true, false); }},
    {"lin_640", []() { return test_linear("lin_640 49152x640x640", 49152, 640, 640, true, true, false); }},
    {"lin_big16", []() { return test_linear("lin_big16 16384x1280x1280", 16384, 1280, 1280, true, true, false); }},
    {"lin_320", []() { return test_linear("lin_320 49152x320x1280", 49152, 320, 1280, true, false, false); }},
    {"lin_ff", []() { return test_linear("lin_ff 8192x10240x1280", 8192, 10240, 1280, true, false, false); }},
    {"mainloop_n256", []() { return test_linear("mainloop_n256 18944x256x16384 (1 tile/pair, pure k-loop)", 18944, 256, 16384, false, false, false); }},
    {"mainloop_n128", []() { return test_linear("mainloop_n128 18944x128x16384 (1 tile/pair, pure k-loop)", 18944, 128, 16384, false, false, false); }},
    {"mainloop_n64", []() { return test_linear("mainloop_n64 18944x64x16384 (1 tile/pair, pure k-loop)", 18944, 64, 16384, false, false, false); }},
    {"dgrad_small", []() { return test_dgrad("dgrad_small 128x64x128", 128, 64, 128); }},
    {"dgrad_ragged", []() { return test_dgrad("dgrad_ragged 300x200x320", 300, 200, 320); }},
    {"dgrad_big", []() { return test_dgrad("dgrad_big 8192x1280x1280", 8192, 1280, 1280); }},
    {"wgrad_small", []() { return test_wgrad("wgrad_small 128x128x64 store", 128, 128, 64, 0); }},
    {"wgrad_ragged", []() { return test_wgrad("wgrad_ragged 1000x320x200 store", 1000, 320, 200, 0); }},
    {"wgrad_atomic", []() { return test_wgrad("wgrad_atomic 4096x320x640 splitK", 4096, 320, 640, 1); }},
    {"wgrad_big", []() { return test_wgrad("wgrad_big 8192x1280x1280 store", 8192, 1280, 1280, 0); }},
    {"conv_small", []() { return test_conv("conv_small 1x16x16 64->64 k3", 1, 16, 16, 64, 64, 3, false); }},
    {"conv_w128", []() { return test_conv("conv_w128 1x8x128 64->32 k3", 1, 8, 128, 64, 32, 3, false); }},
    {"conv_extras", []() { return test_conv("conv_extras 2x32x32 128->320 k3 +bias+img+res", 2, 32, 32, 128, 320, 3, true); }},
    {"conv_1x1", []() { return test_conv("conv_1x1 2x32x32 192->64", 2, 32, 32, 192, 64, 1, true); }},
    {"conv_bucket", []() { return test_conv("conv_bucket 1x36x28 64->64 k3", 1, 36, 28, 64, 64, 3, true); }},
    {"conv_cout4", []() { return test_conv("conv_cout4 1x32x32 320->4 k3", 1, 32, 32, 320, 4, 3, true); }},
    {"conv_big", []() { return test_conv("conv_big 4x64x64 640->640 k3", 4, 64, 64, 640, 640, 3, true); }},
    {"conv_sdxl128", []() { return test_conv("conv_sdxl128 2x128x128 320->320 k3", 2, 128, 128, 320, 320, 3, true); }},
    {"conv_t_ragged", []() { return test_conv("conv_t_ragged 2x36x28 64->128 k3 +bias+img+res (transposed path)", 2, 36, 28, 64, 128, 3, true); }},
    {"conv_t_1x1", []() { return test_conv("conv_t_1x1 3x20x50 128->96 k1 (transposed path)", 3, 20, 50, 128, 96, 1, true); }},
    {"conv_vae1024", []() { return test_conv("conv_vae1024 1x1024x1024 128->128 k3", 1, 1024, 1024, 128, 128, 3, true); }},
    {"cwgrad_small", []() { return test_conv_wgrad("cwgrad_small 1x16x16 64->64 k3", 1, 16, 16, 64, 64, 3); }},
    {"cwgrad_mid", []() { return test_conv_wgrad("cwgrad_mid 2x32x32 128->192 k3", 2, 32, 32, 128, 192, 3); }},
    {"cwgrad_bucket", []() { return test_conv_wgrad("cwgrad_bucket 1x36x28 64->64 k3", 1, 36, 28, 64, 64, 3); }},
    {"cwgrad_w128", []() { return test_conv_wgrad("cwgrad_w128 1x16x128 64->128 k3", 1, 16, 128, 64, 128, 3); }},
    {"cwgrad_sdxl128", []() { return test_conv_wgrad("cwgrad_sdxl128 4x128x128 320->320 k3", 4, 128, 128, 320, 320, 3); }},
    {"cwgrad_sdxl128b16", []() { return test_conv_wgrad("cwgrad_sdxl128b16 16x128x128 320->320 k3", 16, 128, 128, 320, 320, 3); }},
    {"cwgrad_sdxl32", []() { return test_conv_wgrad("cwgrad_sdxl32 8x32x32 1280->1280 k3", 8, 32, 32, 1280, 1280, 3); }},
    {"bqk_d64", []() { return test_batched_qk("bqk_d64 B2 H3 256x200 d64", 2, 3, 256, 200, 64); }},
    {"bqk_d40", []() { return test_batched_qk("bqk_d40 B1 H8 128x77 d40", 1, 8, 128, 77, 40); }},
};

int main(int argc, char** argv) {
    if (argc < 2 || !strcmp(argv[1], "list")) {
        for (auto& c : cases) printf("%s\n", c.name);
        return 0;
    }
    int fails = 0, ran = 0;
    for (auto& c : cases) {
        if (!strcmp(argv[1], "all") || !strcmp(argv[1], c.name)) {
            fails += c.fn();
            ran++;
        }
    }
    if (!ran) {
        printf("unknown case %s\n", argv[1]);
        return 4;
    }
    return fails ? 1 : 0;
}
