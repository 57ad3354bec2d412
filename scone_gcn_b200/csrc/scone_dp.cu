// scone_dp.cu — the one exchange of a data-parallel optimizer step, fused with the optimizer over NVLink peer memory:
//
//   grads <- sum over ranks of [grads | nll_sum | count]      (what jax.grad of the batch loss would have produced on one device)
//   Adam update of the weights                                  scone_trajectory_model.py:264-357 (jax.experimental.optimizers.adam)
//
// in ONE kernel per rank instead of an NCCL all-reduce followed by the Adam kernel.  Every rank owns an exchange buffer
// [2 parities][world slots][blocks][slice + header] that its peers map through CUDA IPC (one process per GPU, one node, NVSwitch).
// A block of the kernel owns a slice of the flat gradient vector:
//   1. PUSH: it stores its slice of the local gradients (+ a 4-float header {nll_sum, count, overflow flag}: every block's chunk is
//      self-contained) into slot `rank` of EVERY rank's buffer — plain stores over NVLink, nothing waits on them;
//   2. SIGNAL: __threadfence_system, then a release store of the step's sequence number into that (rank, block) flag of every peer;
//   3. WAIT: acquire-polls its own flags until all `world` ranks' chunks of this block have landed (bounded: a rank that never
//      arrives sets the error flag after ~10 s instead of hanging the GPU);
//   4. REDUCE + ADAM: sums the `world` chunks IN RANK ORDER out of local memory (every rank adds the same numbers in the same order:
//      bit-identical weights on all ranks), writes the summed gradients back (read_grads sees the all-reduced buffer, as with NCCL)
//      and applies the Adam step to its slice.  If any rank's micro-batch overflowed a capacity, every rank skips the update.
// Two parities: a rank can be at most one step ahead of its slowest peer (it cannot leave step k + 1 before that peer signalled
// k + 1, which it does after it finished reading step k), so step k + 1 never overwrites chunks step k is still reading.
#include <cstring>
#include "common.cuh"

namespace {

constexpr int kDpMaxWorld = 8;
constexpr int kDpThreads = 256;
constexpr int kDpHeader = 4;                       // floats per chunk header: nll_sum, count, overflow, pad
constexpr long long kDpTimeoutClocks = 20000000000ll;      // ~10 s at 2 GHz

struct DpPeers {
    float* data[kDpMaxWorld];
    unsigned* flags[kDpMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kDpThreads) dp_allreduce_adam_kernel(DpPeers peers, int rank, int world, int per, int chunk, unsigned seq,
                                                                       float* __restrict__ grad, const float* __restrict__ scal,
                                                                       long long n_params,
                                                                       float* __restrict__ W, float* __restrict__ m, float* __restrict__ v,
                                                                       const int* __restrict__ overflow, float lr, float wd, float b1,
                                                                       float b2, float eps, float c1, float c2, int* __restrict__ err) {
    const int blk = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    const long long n = n_params + 2;
    const long long lo = (long long)blk * per;
    const int cnt = (int)max(0ll, min((long long)per, n - lo));
    const int par = (int)(seq & 1u);
    const size_t slot = ((size_t)(par * world + rank) * nb + blk) * chunk;           // this rank's chunk of this block, in anybody's buffer
    // 1. push
    const float my_over = (overflow != nullptr && *overflow != 0) ? 1.f : 0.f;
    for (int r = 0; r < world; ++r) {
        float* dst = peers.data[r] + slot;
        if (tid == 0) {
            dst[0] = scal[0];                                   // (a copy of grad[n_params], [n_params + 1] taken before the launch:
            dst[1] = scal[1];                                   //  their owner block overwrites them with the sums in step 4)
            dst[2] = my_over;
        }
        for (int i = tid; i < cnt; i += kDpThreads) dst[kDpHeader + i] = grad[lo + i];
    }
    // 2. signal
    __threadfence_system();
    __syncthreads();
    if (tid < world) st_release_sys(peers.flags[tid] + (size_t)(par * world + rank) * nb + blk, seq);
    // 3. wait
    if (tid < world) {
        const unsigned* f = peers.flags[rank] + (size_t)(par * world + tid) * nb + blk;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != seq) {
            if (clock64() - t0 > kDpTimeoutClocks) {
                *err = 1;
                break;
            }
        }
    }
    __syncthreads();
    // 4. reduce in rank order + Adam
    const float* mine = peers.data[rank] + ((size_t)par * world * nb + blk) * chunk;            // slot 0 of this block; slot r at + r * nb * chunk
    const size_t rstride = (size_t)nb * chunk;
    float nll = 0.f, count = 0.f, over = 0.f;
    for (int r = 0; r < world; ++r) {
        nll += __ldcg(mine + r * rstride + 0);                  // (L2 reads: the chunks were written by other GPUs)
        count += __ldcg(mine + r * rstride + 1);
        over += __ldcg(mine + r * rstride + 2);
    }
    for (int i = tid; i < cnt; i += kDpThreads) {
        float g = 0.f;
        for (int r = 0; r < world; ++r) g += __ldcg(mine + r * rstride + kDpHeader + i);
        const long long gi = lo + i;
        if (gi >= n_params) {                                   // the two scalars at the end of the vector
            grad[gi] = gi == n_params ? nll : count;
            continue;
        }
        grad[gi] = g;
        if (over != 0.f) continue;                              // truncated gradients somewhere: nobody updates (the flag stays set)
        const float gg = g / count + 2.f * wd * W[gi];
        const float mi = (1.f - b1) * gg + b1 * m[gi];
        const float vi = (1.f - b2) * gg * gg + b2 * v[gi];
        m[gi] = mi;
        v[gi] = vi;
        W[gi] = W[gi] - lr * (mi / c1) / (sqrtf(vi / c2) + eps);
    }
}

}  // namespace

struct scone_dp {
    int rank = 0, world = 1, nb = 0, per = 0, chunk = 0;
    int64_t n = 0;                                  // floats of the exchanged vector (n_params + 2)
    void* base = nullptr;                           // local exchange buffer: data, then flags
    size_t data_bytes = 0, flag_bytes = 0;
    void* opened[kDpMaxWorld] = {};
    DpPeers peers = {};
    bool connected = false;
    unsigned seq = 0;
    int* d_err = nullptr;
    float* d_scal = nullptr;                        // {nll_sum, count} of the local gradients, copied before each launch
};

extern "C" int scone_dp_create(int32_t rank, int32_t world, int64_t n_floats, scone_dp** out) {
    SCONE_REQUIRE(out && world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world && n_floats >= 3,
                  "scone_dp_create: rank %d of %d (at most %d ranks of one node), %lld floats", rank, world, kDpMaxWorld, (long long)n_floats);
    scone_dp* d = new scone_dp();
    d->rank = rank;
    d->world = world;
    d->n = n_floats;
    d->nb = (int)std::min<int64_t>((n_floats + kDpThreads - 1) / kDpThreads, 64);
    d->per = (int)((n_floats + d->nb - 1) / d->nb);
    d->chunk = (d->per + kDpHeader + 3) & ~3;
    d->data_bytes = ((size_t)2 * world * d->nb * d->chunk * sizeof(float) + 255) & ~(size_t)255;
    d->flag_bytes = (size_t)2 * world * d->nb * sizeof(unsigned);
    SCONE_CUDA(cudaMalloc(&d->base, d->data_bytes + d->flag_bytes));
    SCONE_CUDA(cudaMemset(d->base, 0, d->data_bytes + d->flag_bytes));
    SCONE_CUDA(cudaMalloc((void**)&d->d_err, sizeof(int)));
    SCONE_CUDA(cudaMemset(d->d_err, 0, sizeof(int)));
    SCONE_CUDA(cudaMalloc((void**)&d->d_scal, 4 * sizeof(float)));
    SCONE_CUDA(cudaDeviceSynchronize());
    d->peers.data[rank] = (float*)d->base;
    d->peers.flags[rank] = (unsigned*)((char*)d->base + d->data_bytes);
    d->connected = world == 1;
    *out = d;
    return 0;
}

extern "C" int32_t scone_dp_handle_bytes(void) { return (int32_t)sizeof(cudaIpcMemHandle_t); }

extern "C" int scone_dp_get_handle(scone_dp* d, void* handle_out) {
    SCONE_REQUIRE(d && handle_out, "scone_dp_get_handle: NULL argument");
    cudaIpcMemHandle_t h;
    SCONE_CUDA(cudaIpcGetMemHandle(&h, d->base));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

// handles: world x scone_dp_handle_bytes() bytes, rank-major (what an all-gather of scone_dp_get_handle produces)
extern "C" int scone_dp_open(scone_dp* d, const void* handles) {
    SCONE_REQUIRE(d && handles, "scone_dp_open: NULL argument");
    for (int r = 0; r < d->world; ++r) {
        if (r == d->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
        SCONE_CUDA(cudaIpcOpenMemHandle(&d->opened[r], h, cudaIpcMemLazyEnablePeerAccess));
        d->peers.data[r] = (float*)d->opened[r];
        d->peers.flags[r] = (unsigned*)((char*)d->opened[r] + d->data_bytes);
    }
    d->connected = true;
    return 0;
}

int scone_dp_allreduce_adam(scone_dp* d, float* W, float* m, float* v, float* grad, int64_t n_params, const int* overflow_dev,
                            int32_t step, float lr, float wd, cudaStream_t st) {
    SCONE_REQUIRE(d && d->connected, "scone_dp: peers not connected (scone_dp_open)");
    SCONE_REQUIRE(n_params + 2 == d->n, "scone_dp: created for %lld floats, model has %lld", (long long)d->n, (long long)(n_params + 2));
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float c1 = 1.f - powf(b1, (float)(step + 1)), c2 = 1.f - powf(b2, (float)(step + 1));
    d->seq += 1;
    SCONE_CUDA(cudaMemcpyAsync(d->d_scal, grad + n_params, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    dp_allreduce_adam_kernel<<<d->nb, kDpThreads, 0, st>>>(d->peers, d->rank, d->world, d->per, d->chunk, d->seq, grad, d->d_scal, (long long)n_params, W, m,
                                                          v, overflow_dev, lr, wd, b1, b2, eps, c1, c2, d->d_err);
    SCONE_LAUNCHED();
    return 0;
}

// synchronises the stream; 3 if a launch timed out waiting for a peer
extern "C" int scone_dp_status(scone_dp* d, void* stream) {
    SCONE_REQUIRE(d, "scone_dp_status: NULL argument");
    int e = 0;
    SCONE_CUDA(cudaMemcpyAsync(&e, d->d_err, sizeof(int), cudaMemcpyDeviceToHost, as_stream(stream)));
    SCONE_CUDA(cudaStreamSynchronize(as_stream(stream)));
    if (e) {
        scone_set_error("scone_dp: a rank did not deliver its gradients within the time-out; weights of that step are invalid");
        return 3;
    }
    return 0;
}

extern "C" void scone_dp_destroy(scone_dp* d) {
    if (!d) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < d->world; ++r)
        if (d->opened[r]) cudaIpcCloseMemHandle(d->opened[r]);
    cudaFree(d->base);
    cudaFree(d->d_err);
    cudaFree(d->d_scal);
    delete d;
}
