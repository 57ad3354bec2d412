// umma_probe.cu — stand-alone check of the tcgen05 building blocks the dense layer kernel uses (run on a B200 under gpurun):
//   1. tcgen05.st.16x128b.x2 at lane bases 32*(warp%4) + 16*h: where do a warp's mma-fragment registers {a0,a1,a2,a3} land in TMEM?
//      (read back with tcgen05.ld.32x32b: thread = lane, registers = columns)
//   2. tcgen05.mma.cta_group::1.kind::tf32 with A from TMEM ([128 lanes][K columns]) and B from shared memory in the no-swizzle
//      K-major canonical layout (core matrix = 8 rows x 16 bytes; LBO = K-direction stride, SBO = N-direction stride), D[128][32] in
//      TMEM, completion through tcgen05.commit -> mbarrier; D read back with tcgen05.ld.16x256b.x4 (mma C-fragment layout).
// Every wait is bounded: a protocol error prints and exits instead of hanging the GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu && ./tools/umma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols));
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void st_16x128b_x2(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// no-swizzle K-major shared-memory descriptor (version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

constexpr int kCols = 64;       // TMEM columns allocated: A at [0, 8), D at [32, 64)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (4u << 17) | (8u << 24);   // f32 acc, tf32 x tf32, K-major, N = 32, M = 128

// out1[lane][col] (128 x 8): TMEM contents after the fragment stores; out2[row][n] (128 x 32): D = A * B^T
__global__ void __launch_bounds__(256) probe_kernel(const float* __restrict__ Bmat /* [32][8] = B[n][k] */, float* __restrict__ out1,
                                                    float* __restrict__ out2, int* __restrict__ status) {
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(128) float s_B[4 * 2 * 32];       // [k-chunk 2][n-group 4][8 rows][4 floats]: LBO = 512 B, SBO = 128 B
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    if (warp == 0) tmem_alloc(&s_tmem, kCols);
    if (tid == 0) mbar_init(&s_bar, 1);
    for (int i = tid; i < 32 * 8; i += 256) {
        const int n = i / 8, k = i % 8;
        s_B[(k / 4) * 128 + (n / 8) * 32 + (n % 8) * 4 + (k % 4)] = Bmat[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the tensor core's async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = s_tmem;
    // ---- 1. fragment stores: warp w owns rows 32*(w%4) + 16*(w/4) .. +15; A[row][k] = row + k / 16 ----
    {
        const int row0 = 32 * (warp & 3) + 16 * (warp >> 2);
        const float a0 = (float)(row0 + g) + tig / 16.f, a1 = (float)(row0 + g + 8) + tig / 16.f;
        const float a2 = (float)(row0 + g) + (tig + 4) / 16.f, a3 = (float)(row0 + g + 8) + (tig + 4) / 16.f;
        st_16x128b_x2(tbase + ((uint32_t)row0 << 16), __float_as_uint(a0), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(a3));
        wait_st();
    }
    fence_before();
    __syncthreads();
    fence_after();
    if (warp < 4) {                                         // thread = lane 32*warp + lane
        uint32_t r[8];
        ld_32x32b_x8(tbase + ((uint32_t)(32 * warp) << 16), r);
        wait_ld();
        for (int c = 0; c < 8; ++c) out1[(32 * warp + lane) * 8 + c] = __uint_as_float(r[c]);
    }
    fence_before();
    __syncthreads();
    fence_after();
    // ---- 2. D[128][32] = A[128][8] * B[32][8]^T ----
    if (warp == 0 && lane == 0) {
        const uint64_t desc = make_desc(smem_u32(s_B), 512, 128);
        umma_tf32_ts(tbase + 32, tbase, desc, kIdesc, 0u);
        umma_commit(&s_bar);
    }
    {
        int spins = 0;
        while (!mbar_try_wait(&s_bar, 0)) {
            if (++spins > 2000000) {
                if (tid == 0) *status = 1;
                break;
            }
        }
    }
    fence_after();
    {
        const int row0 = 32 * (warp & 3) + 16 * (warp >> 2);
        uint32_t r[16];
        ld_16x256b_x4(tbase + 32 + ((uint32_t)row0 << 16), r);
        wait_ld();
        for (int j = 0; j < 4; ++j) {                       // n-tile j: r[4j..4j+3] = (g, 2tig), (g, 2tig+1), (g+8, 2tig), (g+8, 2tig+1)
            out2[(row0 + g) * 32 + 8 * j + 2 * tig] = __uint_as_float(r[4 * j + 0]);
            out2[(row0 + g) * 32 + 8 * j + 2 * tig + 1] = __uint_as_float(r[4 * j + 1]);
            out2[(row0 + g + 8) * 32 + 8 * j + 2 * tig] = __uint_as_float(r[4 * j + 2]);
            out2[(row0 + g + 8) * 32 + 8 * j + 2 * tig + 1] = __uint_as_float(r[4 * j + 3]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, kCols);
}

int main() {
    std::vector<float> hB(32 * 8);
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 8; ++k) hB[n * 8 + k] = (float)((n * 3 + k * 5) % 7 - 3) * 0.25f;       // exactly representable in tf32
    float *dB, *d1, *d2;
    int* dst;
    CK(cudaMalloc(&dB, hB.size() * 4));
    CK(cudaMalloc(&d1, 128 * 8 * 4));
    CK(cudaMalloc(&d2, 128 * 32 * 4));
    CK(cudaMalloc(&dst, 4));
    CK(cudaMemset(d1, 0xff, 128 * 8 * 4));
    CK(cudaMemset(d2, 0xff, 128 * 32 * 4));
    CK(cudaMemset(dst, 0, 4));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    probe_kernel<<<1, 256>>>(dB, d1, d2, dst);
    CK(cudaDeviceSynchronize());
    std::vector<float> o1(128 * 8), o2(128 * 32);
    int st = 0;
    CK(cudaMemcpy(o1.data(), d1, o1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o2.data(), d2, o2.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost));
    int bad1 = 0, bad2 = 0;
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 8; ++k) {
            const float want = (float)r + k / 16.f;
            if (o1[r * 8 + k] != want && bad1++ < 8) printf("  st layout: TMEM[lane %d][col %d] = %g, expected %g\n", r, k, o1[r * 8 + k], want);
        }
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 32; ++n) {
            double want = 0;
            for (int k = 0; k < 8; ++k) want += ((double)r + k / 16.0) * hB[n * 8 + k];
            if (std::fabs(o2[r * 32 + n] - want) > 1e-3 * (1 + std::fabs(want)) && bad2++ < 8)
                printf("  mma: D[%d][%d] = %g, expected %g\n", r, n, o2[r * 32 + n], want);
        }
    printf("umma_probe: mbarrier %s; fragment-store layout mismatches %d / 1024; mma mismatches %d / 4096\n", st ? "TIMED OUT" : "ok", bad1, bad2);
    return (st || bad1 || bad2) ? 2 : 0;
}
