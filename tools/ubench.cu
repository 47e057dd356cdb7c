// Pipe-rate micro-benchmarks that decide the softmax design of the attention kernels:
//   mufu_ex2      ex2.approx.ftz.f32                      (MUFU pipe)
//   mufu_bf16x2   ex2.approx.ftz.bf16x2                   (ptxas emits TWO MUFU.EX2.BF16 — checked with cuobjdump)
//   ffma          fma.rn.f32                              (FMA pipe, one element per lane-op)
//   ffma2         fma.rn.f32x2                            (FMA pipe, two elements per lane-op)
//   poly_exp2     Cody-Waite + degree-3 polynomial exp2 on the FMA/ALU pipes (packed f32x2), no MUFU
//   mix           3 of 4 elements on MUFU, 1 of 4 on the polynomial
// Output: elements per clock per SM (148 CTAs x 512 threads, 8 independent chains per thread).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048
#define CH 8

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t ex2bf2(uint32_t x) {
    uint32_t y;
    asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return (static_cast<unsigned long long>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
}
// 2^x for x <= 0 (clamped at -126), two elements: round-to-nearest split x = n + f, f in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (rel. error ~1e-4: below bf16 rounding), exponent patched in with integer adds
__device__ __forceinline__ unsigned long long poly_exp2_x2(unsigned long long x) {
    const unsigned long long MAGIC = pack2(12582912.f, 12582912.f);  // 1.5 * 2^23
    const unsigned long long NMAGIC = pack2(-12582912.f, -12582912.f);
    const unsigned long long ONE = pack2(1.f, 1.f), NEG1 = pack2(-1.f, -1.f);
    const unsigned long long C1 = pack2(0.695146143436431885f, 0.695146143436431885f);
    const unsigned long long C2 = pack2(0.227564394474029541f, 0.227564394474029541f);
    const unsigned long long C3 = pack2(0.077119089663028717f, 0.077119089663028717f);
    unsigned long long t = fadd2(x, MAGIC);          // low mantissa bits of each half now hold round(x)
    unsigned long long n = fadd2(t, NMAGIC);         // round(x) as float
    unsigned long long f = ffma2(n, NEG1, x);        // x - round(x)
    unsigned long long p = ffma2(C3, f, C2);
    p = ffma2(p, f, C1);
    p = ffma2(p, f, ONE);
    uint32_t lo = static_cast<uint32_t>(p) + (static_cast<uint32_t>(t) << 23);
    uint32_t hi = static_cast<uint32_t>(p >> 32) + (static_cast<uint32_t>(t >> 32) << 23);
    return (static_cast<unsigned long long>(hi) << 32) | lo;
}

template <int MODE>
__global__ void __launch_bounds__(512) bench(float* out, long long* cycles, float seed) {
    float v[CH];
    unsigned long long w[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        v[i] = seed * (threadIdx.x + 1) * (i + 1) * -1e-3f;
        w[i] = pack2(v[i], v[i] * 0.5f);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) v[i] = ex2f(v[i]) - 1.5f;
            else if (MODE == 1) w[i] = (w[i] & 0xffffffff00000000ull) | ex2bf2(static_cast<uint32_t>(w[i]) | 0x80008000u);
            else if (MODE == 2) v[i] = fmaf(v[i], 0.999f, -0.001f);
            else if (MODE == 3) w[i] = ffma2(w[i], pack2(0.999f, 0.999f), pack2(-0.001f, -0.001f));
            else if (MODE == 4) w[i] = fadd2(poly_exp2_x2(w[i]), pack2(-1.5f, -1.5f));
            else if (MODE == 5) {  // per 4 elements: 3 MUFU + ... modelled as: chains 0..5 MUFU pairs, 6..7 polynomial pairs
                if (i < 6) v[i] = ex2f(v[i]) - 1.5f;
                else w[i] = fadd2(poly_exp2_x2(w[i]), pack2(-1.5f, -1.5f));
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) acc += v[i] + __uint_as_float(static_cast<uint32_t>(w[i])) + __uint_as_float(static_cast<uint32_t>(w[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, double elems_per_thread_iter) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    bench<MODE><<<148, 512>>>(out, cyc, 1.0f);
    bench<MODE><<<148, 512>>>(out, cyc, 1.0f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("%-12s %s  %.0f clk  -> %.2f elements/clk/SM\n", name, cudaGetErrorString(e), avg,
           512.0 * ITERS * elems_per_thread_iter / avg);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    // accuracy of the polynomial against exp2f on [-20, 0]
    run<0>("mufu_ex2", CH);
    run<1>("mufu_bf16x2", CH * 2);
    run<2>("ffma", CH);
    run<3>("ffma2", CH * 2);
    run<4>("poly_exp2", CH * 2);
    run<5>("mix 6m+2p", 6 + 2 * 2);
    return 0;
}
