// scone_kernels.cu — hand-written sm_100a kernels of the SCoNe hot path.
//
//   layer_fwd_kernel   Hout = act(Hin W0 + (S0 Hin) W1 + (S1 Hin) W2)      (trajectory_experiments.py:145-149)
//   layer_bwd_kernel   Gprev = ((S_k G) W_k^T summed over k) * act'(Hin),  dW_k = Hin^T (S_k G)
//   layer0_*           the C_in = 1 first layer (input = flows X[E][b])
//   readout_kernel     Bcond(last) @ H_L @ w_out, padded log-softmax, NLL and its gradient
//                      (trajectory_experiments.py:151-152,298-303; scone_trajectory_model.py:46,54)
//   adam_kernel        JAX adam update (scone_trajectory_model.py:300,310)
//
// Design: activations are H[E][b][C] fp32, one edge row = b*C contiguous floats.  A CTA
// owns a tile of TE edges x 128 columns.  Phase 1: every warp owns whole output edge rows and gathers
// the 1 + nnz(S0 row) + nnz(S1 row) neighbour rows with 128-bit coalesced loads in the CSR's fixed
// (ascending) order — no atomics, deterministic — into shared memory.  Phase 2: the three C x C weight
// products, the three-term sum and the activation run as a register-tiled contraction out of shared
// memory, and the result is stored with 128-bit coalesced stores.  Weight gradients are accumulated in
// registers across the tiles of a persistent CTA, written as per-CTA partials and reduced in CTA order.
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTileRows = 128;      // (edge, trajectory) rows per tile
constexpr int kTileCols = 128;      // gathered columns per edge per tile (32 lanes x float4)

template <int ACT>
__device__ __forceinline__ float act_fn(float z) {
    if (ACT == SCONE_ACT_TANH) return tanhf(z);
    if (ACT == SCONE_ACT_LEAKY_RELU) return z >= 0.f ? z : 0.01f * z;
    return fmaxf(z, 0.f);
}
// derivative expressed through the OUTPUT h = act(z) (sign(z) == sign(h) for leaky-relu / relu)
template <int ACT>
__device__ __forceinline__ float dact_fn(float h) {
    if (ACT == SCONE_ACT_TANH) return 1.f - h * h;
    if (ACT == SCONE_ACT_LEAKY_RELU) return h >= 0.f ? 1.f : 0.01f;
    return h > 0.f ? 1.f : 0.f;
}
__device__ __forceinline__ float dact_rt(int act, float h) {
    if (act == SCONE_ACT_TANH) return 1.f - h * h;
    if (act == SCONE_ACT_LEAKY_RELU) return h >= 0.f ? 1.f : 0.01f;
    return h > 0.f ? 1.f : 0.f;
}

__device__ __forceinline__ void fma4(float4& a, float s, const float4& v) {
    a.x = fmaf(s, v.x, a.x);
    a.y = fmaf(s, v.y, a.y);
    a.z = fmaf(s, v.z, a.z);
    a.w = fmaf(s, v.w, a.w);
}

// sum_p coef_p * H[col_p][colofs .. colofs+3] over one CSR row, ascending column order.
__device__ __forceinline__ float4 gather_row4(const float* __restrict__ H, size_t rowlen, int colofs, DevCsr S, int e) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int p = __ldg(S.rowptr + e);
    const int end = __ldg(S.rowptr + e + 1);
    for (; p + 4 <= end; p += 4) {
        int2 c0 = __ldg(S.ent + p), c1 = __ldg(S.ent + p + 1), c2 = __ldg(S.ent + p + 2), c3 = __ldg(S.ent + p + 3);
        float4 v0 = __ldg(reinterpret_cast<const float4*>(H + (size_t)c0.x * rowlen + colofs));
        float4 v1 = __ldg(reinterpret_cast<const float4*>(H + (size_t)c1.x * rowlen + colofs));
        float4 v2 = __ldg(reinterpret_cast<const float4*>(H + (size_t)c2.x * rowlen + colofs));
        float4 v3 = __ldg(reinterpret_cast<const float4*>(H + (size_t)c3.x * rowlen + colofs));
        fma4(acc, __int_as_float(c0.y), v0);
        fma4(acc, __int_as_float(c1.y), v1);
        fma4(acc, __int_as_float(c2.y), v2);
        fma4(acc, __int_as_float(c3.y), v3);
    }
    for (; p < end; ++p) {
        int2 c0 = __ldg(S.ent + p);
        float4 v0 = __ldg(reinterpret_cast<const float4*>(H + (size_t)c0.x * rowlen + colofs));
        fma4(acc, __int_as_float(c0.y), v0);
    }
    return acc;
}

// scalar flavour for the flows X[E][b]
__device__ __forceinline__ float gather_row1(const float* __restrict__ X, int b, int t, DevCsr S, int e) {
    float acc = 0.f;
    int p = __ldg(S.rowptr + e);
    const int end = __ldg(S.rowptr + e + 1);
    for (; p < end; ++p) {
        int2 c0 = __ldg(S.ent + p);
        acc = fmaf(__int_as_float(c0.y), __ldg(X + (size_t)c0.x * b + t), acc);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Register-tiled contraction out[128][NOUT] = T[128][KD] * Wm[KD][NOUT] from shared memory.
// Thread (tx, ty): columns 4*tx .. 4*tx+3, rows ty + i*NRT (i < RT).
// ---------------------------------------------------------------------------------------------
template <int KD, int NOUT, int LDT>
struct TileGemm {
    static constexpr int NTX = NOUT / 4;
    static constexpr int NRT = kThreads / NTX;
    static constexpr int RT = kTileRows / NRT;
    static_assert(NOUT % 4 == 0 && kThreads % NTX == 0 && kTileRows % NRT == 0, "tile shape");
    __device__ __forceinline__ static void run(const float* __restrict__ Ts, const float* __restrict__ Ws, float4 (&acc)[RT]) {
        const int tx = threadIdx.x % NTX, ty = threadIdx.x / NTX;
#pragma unroll
        for (int i = 0; i < RT; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int kq = 0; kq < KD / 4; ++kq) {
            float4 w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const float4*>(Ws + (4 * kq + q) * NOUT + 4 * tx);
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(Ts + (ty + i * NRT) * LDT + 4 * kq);
                fma4(acc[i], a.x, w[0]);
                fma4(acc[i], a.y, w[1]);
                fma4(acc[i], a.z, w[2]);
                fma4(acc[i], a.w, w[3]);
            }
        }
    }
};

// =============================================================================================
// Forward conv layer, C_in, C_out in {8,16,32,64}.
// =============================================================================================
template <int CIN, int COUT, int ACT>
__global__ void __launch_bounds__(kThreads) layer_fwd_kernel(const float* __restrict__ Hin, float* __restrict__ Hout,
                                                            const float* __restrict__ W0, const float* __restrict__ W1,
                                                            const float* __restrict__ W2, DevCsr S0, DevCsr S1, int E, int b) {
    constexpr int TT = kTileCols / CIN;      // trajectories per tile
    constexpr int TE = kTileRows / TT;       // edges per tile
    constexpr int KD = 3 * CIN, LDT = KD + 4;
    using Gemm = TileGemm<KD, COUT, LDT>;
    extern __shared__ __align__(16) float smem[];
    float* Ts = smem;                        // [128][LDT]
    float* Ws = smem + kTileRows * LDT;      // [KD][COUT]

    for (int i = threadIdx.x; i < CIN * COUT; i += kThreads) {
        Ws[i] = W0[i];
        Ws[CIN * COUT + i] = W1[i];
        Ws[2 * CIN * COUT + i] = W2[i];
    }
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const size_t rowlen_in = (size_t)b * CIN, rowlen_out = (size_t)b * COUT;
    const int n_tb = (b + TT - 1) / TT, n_eb = (E + TE - 1) / TE;
    const int jl = (4 * lane) / CIN, cil = (4 * lane) % CIN;

    for (int tile = blockIdx.x; tile < n_tb * n_eb; tile += gridDim.x) {
        const int tb = tile % n_tb, eb = tile / n_tb;
        const int e0 = eb * TE, t0 = tb * TT;
        const int colofs = tb * kTileCols + 4 * lane;
        const bool colok = colofs < (int)rowlen_in;
        __syncthreads();                      // previous tile's contraction done (and Ws visible)
        for (int r = warp; r < TE; r += kWarps) {
            const int e = e0 + r;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
            if (e < E && colok) {
                a0 = __ldg(reinterpret_cast<const float4*>(Hin + (size_t)e * rowlen_in + colofs));
                a1 = gather_row4(Hin, rowlen_in, colofs, S0, e);
                a2 = gather_row4(Hin, rowlen_in, colofs, S1, e);
            }
            float* dst = Ts + (r * TT + jl) * LDT + cil;
            *reinterpret_cast<float4*>(dst) = a0;
            *reinterpret_cast<float4*>(dst + CIN) = a1;
            *reinterpret_cast<float4*>(dst + 2 * CIN) = a2;
        }
        __syncthreads();
        float4 acc[Gemm::RT];
        Gemm::run(Ts, Ws, acc);
        const int tx = threadIdx.x % Gemm::NTX, ty = threadIdx.x / Gemm::NTX;
#pragma unroll
        for (int i = 0; i < Gemm::RT; ++i) {
            const int rho = ty + i * Gemm::NRT;
            const int e = e0 + rho / TT, t = t0 + rho % TT;
            if (e < E && t < b) {
                float4 o;
                o.x = act_fn<ACT>(acc[i].x);
                o.y = act_fn<ACT>(acc[i].y);
                o.z = act_fn<ACT>(acc[i].z);
                o.w = act_fn<ACT>(acc[i].w);
                *reinterpret_cast<float4*>(Hout + (size_t)e * rowlen_out + (size_t)t * COUT + 4 * tx) = o;
            }
        }
    }
}

// =============================================================================================
// Backward conv layer.  A_k = S_k G (k = 1,2; A_0 = G), Gprev = (sum_k A_k W_k^T) * act'(Hin),
// dW_k[ci][co] = sum_rows Hin[row][ci] * A_k[row][co]   (S_k symmetric: (S_k Hin)^T G == Hin^T (S_k G)).
// =============================================================================================
template <int CIN, int COUT>
struct BwdShape {
    static constexpr int TT = kTileCols / COUT, TE = kTileRows / TT;
    static constexpr int KD = 3 * COUT, LDA = KD + 4, LDH = CIN + 4;
    static constexpr int UNITS = COUT * CIN / 4;                       // (co, ci-quad) pairs
    static constexpr int UPT = UNITS >= kThreads ? UNITS / kThreads : 1;  // units per thread
    static constexpr int RS = UNITS >= kThreads ? 1 : kThreads / UNITS;   // row split
    static constexpr size_t smem_floats = (size_t)kTileRows * LDA + (size_t)kTileRows * LDH + (size_t)KD * CIN;
    static constexpr int DW = 3 * CIN * COUT;
};

template <int CIN, int COUT, int ACT, bool WRITE_GPREV>
__global__ void __launch_bounds__(kThreads) layer_bwd_kernel(const float* __restrict__ G, const float* __restrict__ Hin,
                                                            float* __restrict__ Gprev, const float* __restrict__ W0,
                                                            const float* __restrict__ W1, const float* __restrict__ W2,
                                                            float* __restrict__ dw_partial, DevCsr S0, DevCsr S1, int E, int b) {
    using Sh = BwdShape<CIN, COUT>;
    constexpr int TT = Sh::TT, TE = Sh::TE, KD = Sh::KD, LDA = Sh::LDA, LDH = Sh::LDH;
    using Gemm = TileGemm<KD, CIN, LDA>;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                          // [128][LDA]   rows = (edge, traj), cols = k*COUT + co
    float* Hs = As + kTileRows * LDA;          // [128][LDH]
    float* Wt = Hs + kTileRows * LDH;          // [KD][CIN]    Wt[k*COUT+co][ci] = W_k[ci][co]

    for (int i = threadIdx.x; i < CIN * COUT; i += kThreads) {
        const int ci = i / COUT, co = i % COUT;
        Wt[(co)*CIN + ci] = W0[i];
        Wt[(COUT + co) * CIN + ci] = W1[i];
        Wt[(2 * COUT + co) * CIN + ci] = W2[i];
    }
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const size_t rowlen_g = (size_t)b * COUT, rowlen_h = (size_t)b * CIN;
    const int n_tb = (b + TT - 1) / TT, n_eb = (E + TE - 1) / TE;
    const int jl = (4 * lane) / COUT, col = (4 * lane) % COUT;

    float dw[Sh::UPT][3][4];
#pragma unroll
    for (int u = 0; u < Sh::UPT; ++u)
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) dw[u][k][q] = 0.f;

    for (int tile = blockIdx.x; tile < n_tb * n_eb; tile += gridDim.x) {
        const int tb = tile % n_tb, eb = tile / n_tb;
        const int e0 = eb * TE, t0 = tb * TT;
        const int colofs = tb * kTileCols + 4 * lane;
        const bool colok = colofs < (int)rowlen_g;
        __syncthreads();
        // phase 1a: gather A tile
        for (int r = warp; r < TE; r += kWarps) {
            const int e = e0 + r;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
            if (e < E && colok) {
                a0 = __ldg(reinterpret_cast<const float4*>(G + (size_t)e * rowlen_g + colofs));
                a1 = gather_row4(G, rowlen_g, colofs, S0, e);
                a2 = gather_row4(G, rowlen_g, colofs, S1, e);
            }
            float* dst = As + (r * TT + jl) * LDA + col;
            *reinterpret_cast<float4*>(dst) = a0;
            *reinterpret_cast<float4*>(dst + COUT) = a1;
            *reinterpret_cast<float4*>(dst + 2 * COUT) = a2;
        }
        // phase 1b: Hin tile (rows rho = r*TT + j, CIN columns)
        for (int idx = threadIdx.x; idx < kTileRows * (CIN / 4); idx += kThreads) {
            const int rho = idx / (CIN / 4), c4 = idx % (CIN / 4);
            const int e = e0 + rho / TT, t = t0 + rho % TT;
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < E && t < b) h = __ldg(reinterpret_cast<const float4*>(Hin + (size_t)e * rowlen_h + (size_t)t * CIN + 4 * c4));
            *reinterpret_cast<float4*>(Hs + rho * LDH + 4 * c4) = h;
        }
        __syncthreads();
        // phase 2a: Gprev tile
        if (WRITE_GPREV) {
            float4 acc[Gemm::RT];
            Gemm::run(As, Wt, acc);
            const int tx = threadIdx.x % Gemm::NTX, ty = threadIdx.x / Gemm::NTX;
#pragma unroll
            for (int i = 0; i < Gemm::RT; ++i) {
                const int rho = ty + i * Gemm::NRT;
                const int e = e0 + rho / TT, t = t0 + rho % TT;
                if (e < E && t < b) {
                    const float4 h = *reinterpret_cast<const float4*>(Hs + rho * LDH + 4 * tx);
                    float4 o;
                    o.x = acc[i].x * dact_fn<ACT>(h.x);
                    o.y = acc[i].y * dact_fn<ACT>(h.y);
                    o.z = acc[i].z * dact_fn<ACT>(h.z);
                    o.w = acc[i].w * dact_fn<ACT>(h.w);
                    *reinterpret_cast<float4*>(Gprev + (size_t)e * rowlen_h + (size_t)t * CIN + 4 * tx) = o;
                }
            }
        }
        // phase 2b: weight-gradient accumulation (rows of zero-padded tiles contribute 0)
#pragma unroll
        for (int u = 0; u < Sh::UPT; ++u) {
            const int unit = (Sh::UNITS >= kThreads) ? (int)threadIdx.x + u * kThreads : (int)threadIdx.x % Sh::UNITS;
            const int split = (Sh::UNITS >= kThreads) ? 0 : (int)threadIdx.x / Sh::UNITS;
            const int co = unit % COUT, ciq = unit / COUT;
#pragma unroll 4
            for (int rho = split; rho < kTileRows; rho += Sh::RS) {
                const float4 h = *reinterpret_cast<const float4*>(Hs + rho * LDH + 4 * ciq);
                const float a0 = As[rho * LDA + co], a1 = As[rho * LDA + COUT + co], a2 = As[rho * LDA + 2 * COUT + co];
                dw[u][0][0] = fmaf(h.x, a0, dw[u][0][0]); dw[u][0][1] = fmaf(h.y, a0, dw[u][0][1]);
                dw[u][0][2] = fmaf(h.z, a0, dw[u][0][2]); dw[u][0][3] = fmaf(h.w, a0, dw[u][0][3]);
                dw[u][1][0] = fmaf(h.x, a1, dw[u][1][0]); dw[u][1][1] = fmaf(h.y, a1, dw[u][1][1]);
                dw[u][1][2] = fmaf(h.z, a1, dw[u][1][2]); dw[u][1][3] = fmaf(h.w, a1, dw[u][1][3]);
                dw[u][2][0] = fmaf(h.x, a2, dw[u][2][0]); dw[u][2][1] = fmaf(h.y, a2, dw[u][2][1]);
                dw[u][2][2] = fmaf(h.z, a2, dw[u][2][2]); dw[u][2][3] = fmaf(h.w, a2, dw[u][2][3]);
            }
        }
    }
    // per-CTA partial: dw_partial[cta][k][ci][co]; row splits are combined in split order via smem
    __syncthreads();
    float* red = smem;   // reuse: [RS][DW]
    float* outp = dw_partial + (size_t)blockIdx.x * Sh::DW;
#pragma unroll
    for (int u = 0; u < Sh::UPT; ++u) {
        const int unit = (Sh::UNITS >= kThreads) ? (int)threadIdx.x + u * kThreads : (int)threadIdx.x % Sh::UNITS;
        const int split = (Sh::UNITS >= kThreads) ? 0 : (int)threadIdx.x / Sh::UNITS;
        const int co = unit % COUT, ciq = unit / COUT;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int o = (k * CIN + 4 * ciq + q) * COUT + co;
                if (Sh::RS == 1) outp[o] = dw[u][k][q];
                else red[split * Sh::DW + o] = dw[u][k][q];
            }
    }
    if (Sh::RS > 1) {
        __syncthreads();
        for (int o = threadIdx.x; o < Sh::DW; o += kThreads) {
            float s = 0.f;
            for (int sp = 0; sp < Sh::RS; ++sp) s += red[sp * Sh::DW + o];
            outp[o] = s;
        }
    }
}

// out[i] (+)= sum over parts p (ascending) of partial[p][i]
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nparts, int n, float* __restrict__ out,
                                       int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * n + i];
    out[i] = accumulate ? out[i] + s : s;
}

// =============================================================================================
// First layer (C_in = 1): input X[E][b].
// =============================================================================================
constexpr int kL0Edges = 32;     // edges per tile; 32 trajectories per tile (lane = trajectory)

__device__ __forceinline__ void layer0_gather(const float* __restrict__ X, DevCsr S0, DevCsr S1, int E, int b, int e0, int t0,
                                              float* ts /* [3][kL0Edges][32] */) {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int t = t0 + lane;
    for (int r = warp; r < kL0Edges; r += kWarps) {
        const int e = e0 + r;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        if (e < E && t < b) {
            a0 = __ldg(X + (size_t)e * b + t);
            a1 = gather_row1(X, b, t, S0, e);
            a2 = gather_row1(X, b, t, S1, e);
        }
        ts[(0 * kL0Edges + r) * 32 + lane] = a0;
        ts[(1 * kL0Edges + r) * 32 + lane] = a1;
        ts[(2 * kL0Edges + r) * 32 + lane] = a2;
    }
}

template <int COUT, int ACT>
__global__ void __launch_bounds__(kThreads) layer0_fwd_kernel(const float* __restrict__ X, float* __restrict__ Hout,
                                                             const float* __restrict__ W0, const float* __restrict__ W1,
                                                             const float* __restrict__ W2, DevCsr S0, DevCsr S1, int E, int b) {
    __shared__ float ts[3 * kL0Edges * 32];
    __shared__ __align__(16) float ws[3 * COUT];
    for (int i = threadIdx.x; i < COUT; i += kThreads) {
        ws[i] = W0[i];
        ws[COUT + i] = W1[i];
        ws[2 * COUT + i] = W2[i];
    }
    constexpr int C4 = COUT / 4;
    const int n_tb = (b + 31) / 32, n_eb = (E + kL0Edges - 1) / kL0Edges;
    for (int tile = blockIdx.x; tile < n_tb * n_eb; tile += gridDim.x) {
        const int tb = tile % n_tb, eb = tile / n_tb;
        const int e0 = eb * kL0Edges, t0 = tb * 32;
        __syncthreads();
        layer0_gather(X, S0, S1, E, b, e0, t0, ts);
        __syncthreads();
        for (int idx = threadIdx.x; idx < kL0Edges * 32 * C4; idx += kThreads) {
            const int c4 = idx % C4, j = (idx / C4) % 32, r = idx / (C4 * 32);
            const int e = e0 + r, t = t0 + j;
            if (e < E && t < b) {
                const float a0 = ts[(0 * kL0Edges + r) * 32 + j], a1 = ts[(1 * kL0Edges + r) * 32 + j],
                            a2 = ts[(2 * kL0Edges + r) * 32 + j];
                const float4 w0 = *reinterpret_cast<const float4*>(ws + 4 * c4);
                const float4 w1 = *reinterpret_cast<const float4*>(ws + COUT + 4 * c4);
                const float4 w2 = *reinterpret_cast<const float4*>(ws + 2 * COUT + 4 * c4);
                float4 o;
                o.x = act_fn<ACT>(fmaf(a2, w2.x, fmaf(a1, w1.x, a0 * w0.x)));
                o.y = act_fn<ACT>(fmaf(a2, w2.y, fmaf(a1, w1.y, a0 * w0.y)));
                o.z = act_fn<ACT>(fmaf(a2, w2.z, fmaf(a1, w1.z, a0 * w0.z)));
                o.w = act_fn<ACT>(fmaf(a2, w2.w, fmaf(a1, w1.w, a0 * w0.w)));
                *reinterpret_cast<float4*>(Hout + ((size_t)e * b + t) * COUT + 4 * c4) = o;
            }
        }
    }
}

// dW_k[0][co] = sum_{e,t} (S_k X)[e][t] * G0[e][t][co]
template <int COUT>
__global__ void __launch_bounds__(kThreads) layer0_bwd_kernel(const float* __restrict__ X, const float* __restrict__ G0,
                                                             float* __restrict__ dw_partial, DevCsr S0, DevCsr S1, int E, int b) {
    __shared__ float ts[3 * kL0Edges * 32];
    __shared__ float red[kThreads * 12];
    constexpr int C4 = COUT / 4;
    static_assert(kThreads % C4 == 0, "thread->channel-quad mapping must be tile independent");
    float acc[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const int n_tb = (b + 31) / 32, n_eb = (E + kL0Edges - 1) / kL0Edges;
    for (int tile = blockIdx.x; tile < n_tb * n_eb; tile += gridDim.x) {
        const int tb = tile % n_tb, eb = tile / n_tb;
        const int e0 = eb * kL0Edges, t0 = tb * 32;
        __syncthreads();
        layer0_gather(X, S0, S1, E, b, e0, t0, ts);
        __syncthreads();
        for (int idx = threadIdx.x; idx < kL0Edges * 32 * C4; idx += kThreads) {
            const int c4 = idx % C4, j = (idx / C4) % 32, r = idx / (C4 * 32);
            const int e = e0 + r, t = t0 + j;
            if (e < E && t < b) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(G0 + ((size_t)e * b + t) * COUT + 4 * c4));
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float a = ts[(k * kL0Edges + r) * 32 + j];
                    acc[k][0] = fmaf(a, g.x, acc[k][0]);
                    acc[k][1] = fmaf(a, g.y, acc[k][1]);
                    acc[k][2] = fmaf(a, g.z, acc[k][2]);
                    acc[k][3] = fmaf(a, g.w, acc[k][3]);
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) red[threadIdx.x * 12 + k * 4 + q] = acc[k][q];
    __syncthreads();
    // thread o < 3*COUT sums its (k, co) over the kThreads / C4 threads that own channel quad co/4, ascending
    for (int o = threadIdx.x; o < 3 * COUT; o += kThreads) {
        const int k = o / COUT, co = o % COUT, c4 = co / 4, q = co % 4;
        float s = 0.f;
        for (int th = c4; th < kThreads; th += C4) s += red[th * 12 + k * 4 + q];
        dw_partial[(size_t)blockIdx.x * 3 * COUT + o] = s;
    }
}

// X[E][b] from sparse flows; one warp per trajectory (X pre-zeroed).
__global__ void flows_to_dense_kernel(const int32_t* __restrict__ traj_ptr, const int32_t* __restrict__ flow_edge,
                                      const float* __restrict__ flow_val, float* __restrict__ X, int E, int b) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (t >= b) return;
    for (int p = traj_ptr[t] + lane; p < traj_ptr[t + 1]; p += 32) {
        const int e = flow_edge[p];
        if (e >= 0 && e < E) X[(size_t)e * b + t] = flow_val[p];
    }
}

// =============================================================================================
// Readout: one warp per trajectory.
// =============================================================================================
constexpr int kReadoutMaxD = 128, kReadoutMaxCper = 4;   // D <= 128, C <= 128

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(128) readout_kernel(const float* __restrict__ HL, const float* __restrict__ wout,
                                                     const int32_t* __restrict__ last_nodes, const int32_t* __restrict__ nbrhoods,
                                                     const int32_t* __restrict__ inc_ptr, const int2* __restrict__ inc_ent,
                                                     float* __restrict__ logprobs, const int32_t* __restrict__ target_idx,
                                                     const float* __restrict__ mask, float scale, float* __restrict__ GL,
                                                     float* __restrict__ partial /* [b][C+2] */, int act, int N, int D, int b, int C) {
    __shared__ float s_logit[4][kReadoutMaxD];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int t = blockIdx.x * 4 + warp;
    if (t >= b) return;
    float* logit = s_logit[warp];
    const int last = last_nodes[t];
    const bool last_ok = last >= 0 && last < N;
    float w[kReadoutMaxCper];
#pragma unroll
    for (int q = 0; q < kReadoutMaxCper; ++q) w[q] = (lane + 32 * q < C) ? wout[lane + 32 * q] : 0.f;

    for (int j = 0; j < D; ++j) {
        const int nbr = last_ok ? nbrhoods[(size_t)last * D + j] : -1;
        float l = 0.f;                                     // padded slot: row -1 of B1_jax = zeros -> logit 0
        if (nbr >= 0) {
            float part = 0.f;
            float z[kReadoutMaxCper] = {0.f, 0.f, 0.f, 0.f};
            for (int p = inc_ptr[nbr]; p < inc_ptr[nbr + 1]; ++p) {
                const int2 es = inc_ent[p];
                const float* row = HL + ((size_t)es.x * b + t) * C;
#pragma unroll
                for (int q = 0; q < kReadoutMaxCper; ++q)
                    if (lane + 32 * q < C) z[q] = fmaf(__int_as_float(es.y), row[lane + 32 * q], z[q]);
            }
#pragma unroll
            for (int q = 0; q < kReadoutMaxCper; ++q) part = fmaf(z[q], w[q], part);
            l = warp_sum(part);
        }
        if (lane == 0) logit[j] = l;
    }
    __syncwarp();
    float mx = -CUDART_INF_F;
    for (int j = lane; j < D; j += 32) mx = fmaxf(mx, logit[j]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int j = lane; j < D; j += 32) se += expf(logit[j] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    for (int j = lane; j < D; j += 32) logprobs[(size_t)t * D + j] = logit[j] - lse;
    if (GL == nullptr) return;

    const float mk = mask[t];
    const int y = target_idx[t];
    float dwl[kReadoutMaxCper] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < D; ++j) {
        const int nbr = last_ok ? nbrhoods[(size_t)last * D + j] : -1;
        if (nbr < 0) continue;
        const float dl = mk * scale * (expf(logit[j] - lse) - (j == y ? 1.f : 0.f));
        for (int p = inc_ptr[nbr]; p < inc_ptr[nbr + 1]; ++p) {
            const int2 es = inc_ent[p];
            const size_t base = ((size_t)es.x * b + t) * C;
            const float sdl = __int_as_float(es.y) * dl;
#pragma unroll
            for (int q = 0; q < kReadoutMaxCper; ++q)
                if (lane + 32 * q < C) {
                    const float h = HL[base + lane + 32 * q];
                    dwl[q] = fmaf(sdl, h, dwl[q]);                                // z[j][c] * dlogit[j]
                    GL[base + lane + 32 * q] += sdl * w[q] * dact_rt(act, h);      // same lane, sequential: no race
                }
        }
    }
    float* pt = partial + (size_t)t * (C + 2);
#pragma unroll
    for (int q = 0; q < kReadoutMaxCper; ++q)
        if (lane + 32 * q < C) pt[lane + 32 * q] = dwl[q];
    if (lane == 0) {
        pt[C] = (y >= 0 && y < D) ? -mk * (logit[y] - lse) : 0.f;
        pt[C + 1] = mk;
    }
}

// dwout[c] (+)= sum_t partial[t][c]; nll (+)= sum_t partial[t][C]; count (+)= sum_t partial[t][C+1]  (t ascending)
__global__ void readout_reduce_kernel(const float* __restrict__ partial, int b, int C, float* __restrict__ dwout,
                                      float* __restrict__ nll, float* __restrict__ count, int accumulate) {
    const int c = threadIdx.x;
    if (c >= C + 2) return;
    float s = 0.f;
    for (int t = 0; t < b; ++t) s += partial[(size_t)t * (C + 2) + c];
    float* dst = c < C ? dwout + c : (c == C ? nll : count);
    if (dst) *dst = accumulate ? *dst + s : s;
}

__global__ void adam_kernel(float* __restrict__ W, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ gradbuf,
                            long long n, float lr, float wd, float b1, float b2, float eps, float c1, float c2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float count = gradbuf[n + 1];
    const float g = gradbuf[i] / count + 2.f * wd * W[i];
    const float mi = (1.f - b1) * g + b1 * m[i];
    const float vi = (1.f - b2) * g * g + b2 * v[i];
    m[i] = mi;
    v[i] = vi;
    W[i] = W[i] - lr * (mi / c1) / (sqrtf(vi / c2) + eps);
}

int grid_for(const scone_complex* cx, int n_tiles, int ctas_per_sm) {
    int g = cx->num_sms * ctas_per_sm;
    return n_tiles < g ? (n_tiles > 0 ? n_tiles : 1) : g;
}

template <int CIN, int COUT, int ACT>
int launch_fwd(const scone_complex* cx, int b, const float* Hin, const float* W0, const float* W1, const float* W2, float* Hout,
               cudaStream_t st) {
    constexpr int TT = kTileCols / CIN, TE = kTileRows / TT, KD = 3 * CIN, LDT = KD + 4;
    const size_t smem = ((size_t)kTileRows * LDT + (size_t)KD * COUT) * sizeof(float);
    auto kern = layer_fwd_kernel<CIN, COUT, ACT>;
    static bool configured = false;
    if (!configured) {
        SCONE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int occ = 1;
    SCONE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    const int n_tiles = ((b + TT - 1) / TT) * ((cx->E + TE - 1) / TE);
    ScopedProf prof(SCONE_K_LAYER_FWD, st);
    kern<<<grid_for(cx, n_tiles, occ > 0 ? occ : 1), kThreads, smem, st>>>(Hin, Hout, W0, W1, W2, cx->S(0), cx->S(1), cx->E, b);
    SCONE_LAUNCHED();
    return 0;
}

template <int CIN, int COUT>
int dispatch_fwd_act(const scone_complex* cx, int act, int b, const float* Hin, const float* W0, const float* W1, const float* W2,
                     float* Hout, cudaStream_t st) {
    switch (act) {
        case SCONE_ACT_TANH: return launch_fwd<CIN, COUT, SCONE_ACT_TANH>(cx, b, Hin, W0, W1, W2, Hout, st);
        case SCONE_ACT_LEAKY_RELU: return launch_fwd<CIN, COUT, SCONE_ACT_LEAKY_RELU>(cx, b, Hin, W0, W1, W2, Hout, st);
        case SCONE_ACT_RELU: return launch_fwd<CIN, COUT, SCONE_ACT_RELU>(cx, b, Hin, W0, W1, W2, Hout, st);
    }
    scone_set_error("unknown activation %d", act);
    return 2;
}

constexpr int kBwdMaxCtas = 148 * 4;     // upper bound on persistent CTAs (workspace sizing)

template <int CIN, int COUT, int ACT, bool WG>
int launch_bwd(const scone_complex* cx, int b, const float* G, const float* Hin, const float* W0, const float* W1, const float* W2,
               float* Gprev, float* dW, int accumulate, float* ws, cudaStream_t st) {
    using Sh = BwdShape<CIN, COUT>;
    size_t smem_f = Sh::smem_floats;
    if ((size_t)Sh::RS * Sh::DW > smem_f) smem_f = (size_t)Sh::RS * Sh::DW;
    const size_t smem = smem_f * sizeof(float);
    auto kern = layer_bwd_kernel<CIN, COUT, ACT, WG>;
    static bool configured = false;
    if (!configured) {
        SCONE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int occ = 1;
    SCONE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    if (occ < 1) occ = 1;
    const int n_tiles = ((b + Sh::TT - 1) / Sh::TT) * ((cx->E + Sh::TE - 1) / Sh::TE);
    int grid = grid_for(cx, n_tiles, occ);
    if (grid > kBwdMaxCtas) grid = kBwdMaxCtas;
    ScopedProf prof(SCONE_K_LAYER_BWD, st);
    kern<<<grid, kThreads, smem, st>>>(G, Hin, Gprev, W0, W1, W2, ws, cx->S(0), cx->S(1), cx->E, b);
    SCONE_LAUNCHED();
    reduce_partials_kernel<<<(Sh::DW + 255) / 256, 256, 0, st>>>(ws, grid, Sh::DW, dW, accumulate);
    SCONE_LAUNCHED();
    return 0;
}

template <int CIN, int COUT>
int dispatch_bwd_act(const scone_complex* cx, int act, int b, const float* G, const float* Hin, const float* W0, const float* W1,
                     const float* W2, float* Gprev, float* dW, int accumulate, float* ws, cudaStream_t st) {
#define SCONE_BWD_CASE(A)                                                                                                    \
    case A:                                                                                                                  \
        return Gprev ? launch_bwd<CIN, COUT, A, true>(cx, b, G, Hin, W0, W1, W2, Gprev, dW, accumulate, ws, st)              \
                     : launch_bwd<CIN, COUT, A, false>(cx, b, G, Hin, W0, W1, W2, Gprev, dW, accumulate, ws, st);
    switch (act) {
        SCONE_BWD_CASE(SCONE_ACT_TANH)
        SCONE_BWD_CASE(SCONE_ACT_LEAKY_RELU)
        SCONE_BWD_CASE(SCONE_ACT_RELU)
    }
#undef SCONE_BWD_CASE
    scone_set_error("unknown activation %d", act);
    return 2;
}

bool width_ok(int c) { return c == 8 || c == 16 || c == 32 || c == 64; }

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
#define SCONE_DISPATCH_WIDTHS(FN, ...)                                              \
    do {                                                                            \
        const int key_ = cin * 1000 + cout;                                         \
        switch (key_) {                                                             \
            case 8008: return FN<8, 8>(__VA_ARGS__);                                \
            case 8016: return FN<8, 16>(__VA_ARGS__);                               \
            case 16008: return FN<16, 8>(__VA_ARGS__);                              \
            case 16016: return FN<16, 16>(__VA_ARGS__);                             \
            case 16032: return FN<16, 32>(__VA_ARGS__);                             \
            case 32016: return FN<32, 16>(__VA_ARGS__);                             \
            case 32032: return FN<32, 32>(__VA_ARGS__);                             \
            case 32064: return FN<32, 64>(__VA_ARGS__);                             \
            case 64032: return FN<64, 32>(__VA_ARGS__);                             \
            case 64064: return FN<64, 64>(__VA_ARGS__);                             \
        }                                                                           \
    } while (0)

extern "C" int scone_layer_forward(const scone_complex* cx, int32_t act, int32_t b, int32_t cin, int32_t cout, const float* Hin,
                                   const float* W0, const float* W1, const float* W2, float* Hout, void* stream) {
    SCONE_REQUIRE(cx && Hin && W0 && W1 && W2 && Hout, "scone_layer_forward: NULL argument");
    SCONE_REQUIRE(!cx->host_only, "scone_layer_forward: index-only complex has no device arrays");
    SCONE_REQUIRE(b > 0, "scone_layer_forward: b must be positive");
    if (cin == 1) return scone_layer0_forward(cx, act, b, cout, Hin, W0, W1, W2, Hout, stream);
    SCONE_REQUIRE(width_ok(cin) && width_ok(cout),
                  "scone_layer_forward: hidden widths must be in {8,16,32,64} with |log2 ratio| <= 1 (got %d -> %d)", cin, cout);
    SCONE_DISPATCH_WIDTHS(dispatch_fwd_act, cx, act, b, Hin, W0, W1, W2, Hout, as_stream(stream));
    scone_set_error("scone_layer_forward: unsupported width pair %d -> %d", cin, cout);
    return 2;
}

extern "C" int64_t scone_layer_backward_workspace_bytes(int32_t cin, int32_t cout) {
    if (cin == 1) return scone_layer0_backward_workspace_bytes(cout);
    return (int64_t)kBwdMaxCtas * 3 * cin * cout * sizeof(float);
}

extern "C" int scone_layer_backward(const scone_complex* cx, int32_t act, int32_t b, int32_t cin, int32_t cout, const float* G,
                                    const float* Hin, const float* W0, const float* W1, const float* W2, float* Gprev, float* dW,
                                    int32_t accumulate, void* workspace, void* stream) {
    SCONE_REQUIRE(cx && G && Hin && dW && workspace, "scone_layer_backward: NULL argument");
    SCONE_REQUIRE(!cx->host_only, "scone_layer_backward: index-only complex has no device arrays");
    SCONE_REQUIRE(b > 0, "scone_layer_backward: b must be positive");
    if (cin == 1) return scone_layer0_backward(cx, b, cout, G, Hin, dW, accumulate, workspace, stream);
    SCONE_REQUIRE(W0 && W1 && W2, "scone_layer_backward: NULL weights");
    SCONE_REQUIRE(width_ok(cin) && width_ok(cout),
                  "scone_layer_backward: hidden widths must be in {8,16,32,64} with |log2 ratio| <= 1 (got %d -> %d)", cin, cout);
    SCONE_DISPATCH_WIDTHS(dispatch_bwd_act, cx, act, b, G, Hin, W0, W1, W2, Gprev, dW, accumulate, (float*)workspace,
                          as_stream(stream));
    scone_set_error("scone_layer_backward: unsupported width pair %d -> %d", cin, cout);
    return 2;
}

template <int COUT>
static int launch_l0_fwd(const scone_complex* cx, int act, int b, const float* X, const float* W0, const float* W1, const float* W2,
                         float* Hout, cudaStream_t st) {
    const int n_tiles = ((b + 31) / 32) * ((cx->E + kL0Edges - 1) / kL0Edges);
    const int grid = grid_for(cx, n_tiles, 6);
    ScopedProf prof(SCONE_K_LAYER0_FWD, st);
    switch (act) {
        case SCONE_ACT_TANH:
            layer0_fwd_kernel<COUT, SCONE_ACT_TANH><<<grid, kThreads, 0, st>>>(X, Hout, W0, W1, W2, cx->S(0), cx->S(1), cx->E, b);
            break;
        case SCONE_ACT_LEAKY_RELU:
            layer0_fwd_kernel<COUT, SCONE_ACT_LEAKY_RELU><<<grid, kThreads, 0, st>>>(X, Hout, W0, W1, W2, cx->S(0), cx->S(1), cx->E, b);
            break;
        case SCONE_ACT_RELU:
            layer0_fwd_kernel<COUT, SCONE_ACT_RELU><<<grid, kThreads, 0, st>>>(X, Hout, W0, W1, W2, cx->S(0), cx->S(1), cx->E, b);
            break;
        default: scone_set_error("unknown activation %d", act); return 2;
    }
    SCONE_LAUNCHED();
    return 0;
}

int scone_layer0_forward(const scone_complex* cx, int32_t act, int32_t b, int32_t cout, const float* X, const float* W0,
                         const float* W1, const float* W2, float* Hout, void* stream) {
    cudaStream_t st = as_stream(stream);
    switch (cout) {
        case 8: return launch_l0_fwd<8>(cx, act, b, X, W0, W1, W2, Hout, st);
        case 16: return launch_l0_fwd<16>(cx, act, b, X, W0, W1, W2, Hout, st);
        case 32: return launch_l0_fwd<32>(cx, act, b, X, W0, W1, W2, Hout, st);
        case 64: return launch_l0_fwd<64>(cx, act, b, X, W0, W1, W2, Hout, st);
    }
    scone_set_error("scone_layer_forward: first-layer width must be in {8,16,32,64} (got %d)", cout);
    return 2;
}

constexpr int kL0BwdCtas = 148 * 4;
int64_t scone_layer0_backward_workspace_bytes(int32_t cout) { return (int64_t)kL0BwdCtas * 3 * cout * sizeof(float); }

template <int COUT>
static int launch_l0_bwd(const scone_complex* cx, int b, const float* G, const float* X, float* dW, int accumulate, float* ws,
                         cudaStream_t st) {
    const int n_tiles = ((b + 31) / 32) * ((cx->E + kL0Edges - 1) / kL0Edges);
    int grid = grid_for(cx, n_tiles, 4);
    if (grid > kL0BwdCtas) grid = kL0BwdCtas;
    ScopedProf prof(SCONE_K_LAYER0_BWD, st);
    layer0_bwd_kernel<COUT><<<grid, kThreads, 0, st>>>(X, G, ws, cx->S(0), cx->S(1), cx->E, b);
    SCONE_LAUNCHED();
    reduce_partials_kernel<<<(3 * COUT + 255) / 256, 256, 0, st>>>(ws, grid, 3 * COUT, dW, accumulate);
    SCONE_LAUNCHED();
    return 0;
}

int scone_layer0_backward(const scone_complex* cx, int32_t b, int32_t cout, const float* G, const float* X, float* dW,
                          int32_t accumulate, void* workspace, void* stream) {
    cudaStream_t st = as_stream(stream);
    float* ws = (float*)workspace;
    switch (cout) {
        case 8: return launch_l0_bwd<8>(cx, b, G, X, dW, accumulate, ws, st);
        case 16: return launch_l0_bwd<16>(cx, b, G, X, dW, accumulate, ws, st);
        case 32: return launch_l0_bwd<32>(cx, b, G, X, dW, accumulate, ws, st);
        case 64: return launch_l0_bwd<64>(cx, b, G, X, dW, accumulate, ws, st);
    }
    scone_set_error("scone_layer_backward: first-layer width must be in {8,16,32,64} (got %d)", cout);
    return 2;
}

extern "C" int scone_flows_to_dense(const scone_complex* cx, int32_t b, const int32_t* traj_ptr, const int32_t* flow_edge,
                                    const float* flow_val, float* X, void* stream) {
    SCONE_REQUIRE(cx && traj_ptr && X && b > 0, "scone_flows_to_dense: bad argument");
    SCONE_REQUIRE(!cx->host_only, "scone_flows_to_dense: index-only complex has no device arrays");
    cudaStream_t st = as_stream(stream);
    ScopedProf prof(SCONE_K_OTHER, st);
    SCONE_CUDA(cudaMemsetAsync(X, 0, (size_t)cx->E * b * sizeof(float), st));
    flows_to_dense_kernel<<<(b * 32 + 255) / 256, 256, 0, st>>>(traj_ptr, flow_edge, flow_val, X, cx->E, b);
    SCONE_LAUNCHED();
    return 0;
}

int64_t scone_readout_workspace_bytes(int32_t b, int32_t C) { return (int64_t)b * (C + 2) * sizeof(float); }

int scone_readout_ws(const scone_complex* cx, int32_t act, int32_t b, int32_t C, const float* HL, const float* wout,
                     const int32_t* last_nodes, float* logprobs, const int32_t* target_idx, const float* mask, float scale,
                     float* GL, float* dwout, float* nll_sum, float* count, int32_t accumulate, void* workspace, void* stream) {
    SCONE_REQUIRE(cx && HL && wout && last_nodes && logprobs, "scone_readout: NULL argument");
    SCONE_REQUIRE(!cx->host_only, "scone_readout: index-only complex has no device arrays");
    SCONE_REQUIRE(C >= 1 && C <= 32 * kReadoutMaxCper, "scone_readout: C must be in [1,%d]", 32 * kReadoutMaxCper);
    SCONE_REQUIRE(cx->D <= kReadoutMaxD, "scone_readout: max degree %d exceeds %d", cx->D, kReadoutMaxD);
    cudaStream_t st = as_stream(stream);
    ScopedProf prof(SCONE_K_READOUT, st);
    if (GL) {
        SCONE_REQUIRE(target_idx && mask && workspace, "scone_readout: gradient mode needs target_idx, mask, workspace");
        SCONE_CUDA(cudaMemsetAsync(GL, 0, (size_t)cx->E * b * C * sizeof(float), st));
    }
    readout_kernel<<<(b + 3) / 4, 128, 0, st>>>(HL, wout, last_nodes, cx->d_nbrhoods, cx->d_inc_ptr, cx->d_inc_ent, logprobs,
                                              target_idx, mask, scale, GL, (float*)workspace, act, cx->N, cx->D, b, C);
    SCONE_LAUNCHED();
    if (GL) {
        readout_reduce_kernel<<<1, ((C + 2 + 31) / 32) * 32, 0, st>>>((const float*)workspace, b, C, dwout, nll_sum, count, accumulate);
        SCONE_LAUNCHED();
    }
    return 0;
}

extern "C" int64_t scone_readout_workspace(int32_t b, int32_t C) { return scone_readout_workspace_bytes(b, C); }

extern "C" int scone_readout(const scone_complex* cx, int32_t act, int32_t b, int32_t C, const float* HL, const float* wout,
                             const int32_t* last_nodes, float* logprobs, const int32_t* target_idx, const float* mask, float scale,
                             float* GL, float* dwout, float* nll_sum, float* count, int32_t accumulate, void* workspace,
                             void* stream) {
    return scone_readout_ws(cx, act, b, C, HL, wout, last_nodes, logprobs, target_idx, mask, scale, GL, dwout, nll_sum, count,
                            accumulate, workspace, stream);
}

int scone_adam_launch(float* W, float* m, float* v, const float* gradbuf, int64_t n, int32_t step, float lr, float wd, void* stream) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float c1 = 1.f - powf(b1, (float)(step + 1)), c2 = 1.f - powf(b2, (float)(step + 1));
    adam_kernel<<<(int)((n + 255) / 256), 256, 0, as_stream(stream)>>>(W, m, v, gradbuf, (long long)n, lr, wd, b1, b2, eps, c1, c2);
    SCONE_LAUNCHED();
    return 0;
}
