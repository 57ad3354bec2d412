// microbench.cu — instruction-throughput probes on sm_100a that the fused-layer kernel design rests on:
//   FFMA (3-register), FFMA2 (fma.rn.f32x2), mma.sync m16n8k8 TF32, tanhf vs the library's own tanh.
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
// Run on the GPU box:  ./tools/microbench
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

constexpr int ITERS = 2048;

__global__ void __launch_bounds__(256) k_ffma(float* out, float s) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    float b = s, c = s * 0.5f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, float s) {
    unsigned long long a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(threadIdx.x * 0.001f + i, i * 0.5f);
    const unsigned long long b = pk(s, s), c = pk(s * 0.5f, s * 0.25f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = ffma2(a[i], b, c);
    }
    unsigned long long r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(r ^ (r >> 32)));
}

__global__ void __launch_bounds__(256) k_mma(float* out, float s) {
    float d[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    uint32_t a[4] = {__float_as_uint(s), __float_as_uint(s * 2), __float_as_uint(s * 3), __float_as_uint(s * 4)};
    uint32_t b0 = __float_as_uint(s * 0.5f), b1 = __float_as_uint(s * 0.25f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) r += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// the library's tanh (copy of scone_tanh in csrc/scone_slab.cuh; keep in sync)
__device__ __forceinline__ float scone_tanh(float x) {
    const float ax = fabsf(x);
    const float x2 = x * x;
    float p = fmaf(x2, -6.1490963126e-03f, 2.0973112000e-02f);     // odd minimax polynomial on |x| < 0.55 (tools/fit_tanh.py)
    p = fmaf(p, x2, -5.3824928855e-02f);
    p = fmaf(p, x2, 1.3332274816e-01f);
    p = fmaf(p, x2, -3.3333305752e-01f);
    const float small = fmaf(x * x2, p, x);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.885390082f));   // exp(2|x|)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    const float big = copysignf(fmaf(-2.f, r, 1.f), x);
    return ax < 0.55f ? small : big;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_tanh(float* out, float s) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (threadIdx.x % 64) * 0.01f * s + i * 0.05f;
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = (MODE == 0 ? tanhf(a[i]) : scone_tanh(a[i])) + 0.3f;
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void k_tanh_err(const float* x, float* y0, float* y1, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { y0[i] = tanhf(x[i]); y1[i] = scone_tanh(x[i]); }
}

template <typename K>
static float time_ms(K launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, grid = sms * 8, thr = 256;
    printf("device %s, %d SMs, clock attr %d MHz\n", p.name, sms, clk_khz / 1000);
    float* out; CK(cudaMalloc(&out, (size_t)grid * thr * 4));
    const double lanes = (double)grid * thr;
    float ms;
    ms = time_ms([&] { k_ffma<<<grid, thr>>>(out, 0.999f); });
    printf("FFMA   : %.3f ms  %.1f TFLOP/s  (%.1f FMA lanes/clk/SM at 1965 MHz)\n", ms, 2.0 * lanes * 8 * ITERS / ms / 1e9,
           lanes * 8 * ITERS / (ms * 1e-3) / sms / 1.965e9);
    ms = time_ms([&] { k_ffma2<<<grid, thr>>>(out, 0.999f); });
    printf("FFMA2  : %.3f ms  %.1f TFLOP/s  (%.1f FMA lanes/clk/SM at 1965 MHz)\n", ms, 4.0 * lanes * 8 * ITERS / ms / 1e9,
           2.0 * lanes * 8 * ITERS / (ms * 1e-3) / sms / 1.965e9);
    ms = time_ms([&] { k_mma<<<grid, thr>>>(out, 0.5f); });
    {
        const double mmas = lanes / 32 * 4 * ITERS;
        printf("mma.sync m16n8k8 tf32: %.3f ms  %.1f TFLOP/s dense-equivalent  (%.2f mma/clk/SM at 1965 MHz)\n", ms,
               mmas * 2 * 16 * 8 * 8 / ms / 1e9, mmas / (ms * 1e-3) / sms / 1.965e9);
    }
    ms = time_ms([&] { k_tanh<0><<<grid, thr>>>(out, 1.f); });
    printf("tanhf      : %.3f ms  %.2f G tanh/s  (%.2f warp-tanh/clk/SM)\n", ms, lanes * ITERS / ms / 1e6,
           lanes / 32 * ITERS / (ms * 1e-3) / sms / 1.965e9);
    ms = time_ms([&] { k_tanh<1><<<grid, thr>>>(out, 1.f); });
    printf("scone_tanh : %.3f ms  %.2f G tanh/s  (%.2f warp-tanh/clk/SM)\n", ms, lanes * ITERS / ms / 1e6,
           lanes / 32 * ITERS / (ms * 1e-3) / sms / 1.965e9);
    // accuracy of scone_tanh against double tanh
    const int n = 1 << 22;
    std::vector<float> hx(n), h0(n), h1(n);
    for (int i = 0; i < n; ++i) {
        const double u = (double)i / n;            // log-spaced magnitudes 1e-6 .. 20, both signs
        hx[i] = (float)((i & 1 ? -1.0 : 1.0) * pow(10.0, -6.0 + 7.3 * u));
    }
    float *dx, *d0, *d1;
    CK(cudaMalloc(&dx, n * 4)); CK(cudaMalloc(&d0, n * 4)); CK(cudaMalloc(&d1, n * 4));
    CK(cudaMemcpy(dx, hx.data(), n * 4, cudaMemcpyHostToDevice));
    k_tanh_err<<<(n + 255) / 256, 256>>>(dx, d0, d1, n);
    CK(cudaMemcpy(h0.data(), d0, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h1.data(), d1, n * 4, cudaMemcpyDeviceToHost));
    double e0a = 0, e0r = 0, e1a = 0, e1r = 0;
    for (int i = 0; i < n; ++i) {
        const double t = tanh((double)hx[i]);
        e0a = fmax(e0a, fabs(h0[i] - t)); e0r = fmax(e0r, fabs(h0[i] - t) / fabs(t));
        e1a = fmax(e1a, fabs(h1[i] - t)); e1r = fmax(e1r, fabs(h1[i] - t) / fabs(t));
    }
    printf("tanh error vs double: tanhf abs %.3g rel %.3g | scone_tanh abs %.3g rel %.3g\n", e0a, e0r, e1a, e1r);
    return 0;
}
