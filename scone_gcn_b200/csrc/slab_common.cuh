// slab_common.cuh — device helpers shared by the slab (dense) and row-list kernels: packed-FMA gathers whose accumulators
// are mma.sync fragments, the 3xTF32 product, the library's tanh.  Included inside an anonymous namespace.
#pragma once
#include "common.cuh"

typedef unsigned long long u64;
constexpr int kSlabWarps = 16;
constexpr int kSlabThreads = kSlabWarps * 32;
constexpr int kRowsWarps = 16;            // row-list kernels (measured: 20 warps at <= 96 registers is 9 % slower than 16 at 128)
constexpr int kRowsThreads = kRowsWarps * 32;
constexpr int kRowsMini = 1;              // consecutive slabs a warp takes per grab of the dynamic counter

__device__ __forceinline__ u64 bcast2(float c) {
    u64 r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(c));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ void ffma2(u64& acc, u64 a, u64 b) {      // acc = (a.x*b.x + acc.x, a.y*b.y + acc.y), each an IEEE fma
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void ldg128(const float* p, u64& lo, u64& hi) {
    asm("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p));
}
// row bitmaps (bit e*b + t mirrors the occupancy flag of row (e, t)); setting a bit is idempotent
__device__ __forceinline__ bool bit_test(const uint32_t* __restrict__ bm, size_t row) { return (__ldg(bm + (row >> 5)) >> (row & 31)) & 1u; }
__device__ __forceinline__ void bit_set(uint32_t* __restrict__ bm, size_t row) {
    uint32_t* w = bm + (row >> 5);
    const uint32_t bit = 1u << (row & 31);
    if (!(*w & bit)) atomicOr(w, bit);
}
// Row ids.  Dense tensors [E][b][C] (pipeline 1) number their rows edge-major, id = e * b + t.  COMPACT tensors (pipelines 2 / 3) are
// numbered TRAJECTORY-major, id = t * E + e: the rows of one trajectory are a few clusters of the locality-ordered edge range, so
// a slab's 16 list rows, their neighbour rows, the bitmap / prefix words of the rank lookups and the X entries of the first layer
// are cache neighbours (edge-major ids put every (e', t) in its own 128-byte line).  toff is what a neighbour's id adds to its
// edge term: t (edge-major) or t * E (trajectory-major).
template <bool TMAJ>
struct RowIds {
    unsigned e, toff;
    __device__ __forceinline__ RowIds(unsigned rid, int b, int E) {
        if (TMAJ) {
            const unsigned t = rid / (unsigned)E;
            toff = t * (unsigned)E;
            e = rid - toff;
        } else {
            e = rid / (unsigned)b;
            toff = rid - e * (unsigned)b;
        }
        this->b = (unsigned)b;
    }
    unsigned b;
    __device__ __forceinline__ unsigned row(unsigned edge) const { return TMAJ ? toff + edge : edge * b + toff; }
    static __device__ __forceinline__ unsigned row_of(unsigned edge, unsigned toff_, int b_) { return TMAJ ? toff_ + edge : edge * (unsigned)b_ + toff_; }
};

// two-level bitmap: bm1 = one summary bit per 32-bit word of bm, set by whoever turns a word non-zero; the compaction and the
// clearing of a sparse bitmap then scan 1/32 of its bytes (compact_summary_kernel, clear_summary_kernel in scone_rows.cu)
__device__ __forceinline__ void bit_set2(uint32_t* __restrict__ bm, uint32_t* __restrict__ bm1, size_t row) {
    const size_t widx = row >> 5;
    uint32_t* w = bm + widx;
    const uint32_t bit = 1u << (row & 31);
    if (!(*w & bit)) {                                   // both atomics are fire-and-forget (RED): nothing waits on their result
        atomicOr(w, bit);
        if (bm1 != nullptr) atomicOr(bm1 + (widx >> 5), 1u << (widx & 31));
    }
}
// compact storage: row r of a tensor lives at index rank(r) = pref[r >> 5] + popc(bm[r >> 5] & bits below r) = its position in
// the compacted row list; returns whether the row exists (bit set)
__device__ __forceinline__ bool rank_lookup(const uint32_t* __restrict__ bm, const uint32_t* __restrict__ pref, unsigned row,
                                            unsigned& idx) {
    const uint32_t w = __ldg(bm + (row >> 5));
    const unsigned sh = row & 31u;
    idx = __ldg(pref + (row >> 5)) + (unsigned)__popc(w & ((1u << sh) - 1u));
    return (w >> sh) & 1u;
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// a = hi + lo for the 3xTF32 product.  hi = a rounded to nearest at 10 mantissa bits (integer add + mask: cvt.rna.tf32 lowers
// to a 4-instruction sequence with an inf/nan guard on sm_100a); lo = a - hi is exact in fp32 and is handed to the tensor
// core as is (the hardware reads the upper 19 bits: |truncation| <= 2^-11 |lo| <= 2^-22 |a|).
__device__ __forceinline__ void split_tf32(float a, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(a) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(a - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tanh with fp32-grade accuracy (max abs error 6e-8, tools/microbench.cu measures it against double): an odd minimax
// polynomial on |x| < 0.55 (tools/fit_tanh.py) and 1 - 2 / (exp(2|x|) + 1) beyond, both from single MUFU ops.
__device__ __forceinline__ float scone_tanh(float x) {
    const float ax = fabsf(x);
    const float x2 = x * x;
    float p = fmaf(x2, -6.1490963126e-03f, 2.0973112000e-02f);
    p = fmaf(p, x2, -5.3824928855e-02f);
    p = fmaf(p, x2, 1.3332274816e-01f);
    p = fmaf(p, x2, -3.3333305752e-01f);
    const float small = fmaf(x * x2, p, x);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.885390082f));   // exp(2|x|)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    const float big = copysignf(fmaf(-2.f, r, 1.f), x);
    return ax < 0.55f ? small : big;
}

__device__ __forceinline__ float scone_tanh_small(float x) {        // |x| < 0.55 only
    const float x2 = x * x;
    float p = fmaf(x2, -6.1490963126e-03f, 2.0973112000e-02f);
    p = fmaf(p, x2, -5.3824928855e-02f);
    p = fmaf(p, x2, 1.3332274816e-01f);
    p = fmaf(p, x2, -3.3333305752e-01f);
    return fmaf(x * x2, p, x);
}

// activation of a lane's NT x 4 accumulator values in place; tanh takes the polynomial-only path when every value of the
// warp is small (warp-uniform branch), which is the common case for the reference's 0.01-scale initialisation
template <int ACT, int NT>
__device__ __forceinline__ void slab_activate(float (&d)[NT][4]) {
    if (ACT == SCONE_ACT_TANH) {
        float m = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) m = fmaxf(m, fabsf(d[nt][q]));
        if (__all_sync(0xffffffffu, m < 0.55f)) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 4; ++q) d[nt][q] = scone_tanh_small(d[nt][q]);
        } else {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 4; ++q) d[nt][q] = scone_tanh(d[nt][q]);
        }
    } else {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float z = d[nt][q];
                d[nt][q] = ACT == SCONE_ACT_LEAKY_RELU ? (z >= 0.f ? z : 0.01f * z) : fmaxf(z, 0.f);
            }
    }
}

// Geometry of a slab for channel width C and TS trajectories per slab.
//   Q   trajectories covered by one warp-wide 128-bit load (512 contiguous bytes of an edge row)
//   NL  load slots per slab (16 rows / Q);  LPE slots per edge;  EPS edges per slab;  KS mma k-steps per term
// Width 32: lane l = (g = l >> 2, tig = l & 3) loads channels 16*(g&1) + 4*tig .. +3 of trajectory g >> 1 in slots 0, 1 and
// the other channel half (lane ^ 4 position) in slots 2, 3; after the lane^4 exchange an even-g lane owns rows from slots
// {0, 1}, an odd-g lane rows from slots {2, 3}; mma k index (step s, kappa) <-> channel 4*kappa + s (kappa < 4) or
// 16 + 4*(kappa-4) + s.   Width 16: lane loads channels 4*tig .. +3 of trajectory g; slot 0 = fragment rows g, slot 1 =
// rows g + 8; (s, kappa) <-> channel 4*kappa + 2*s (kappa < 4) or 4*(kappa-4) + 2*s + 1.
template <int C, int TS>
struct SlabGeom {
    static_assert(C == 16 || C == 32, "slab kernels exist for widths 16 and 32");
    static constexpr int Q = 128 / C, NL = 16 / Q, LPE = TS / Q, EPS = 16 / TS, KS = C / 8;
    static_assert(TS % Q == 0 && 16 % TS == 0 && LPE >= 1, "slab shape");
    __host__ __device__ static constexpr int chan(int s, int kappa) {
        return C == 32 ? (kappa < 4 ? 4 * kappa + s : 16 + 4 * (kappa - 4) + s)
                       : (kappa < 4 ? 4 * kappa + 2 * s : 4 * (kappa - 4) + 2 * s + 1);
    }
};

// Weight fragments in shared memory: Bf[((term * KS + s) * NT + nt) * 32 + lane] = {b0_hi, b1_hi, b0_lo, b1_lo},
// b0 = W_term[chan(s, tig)][nt*8 + g], b1 = W_term[chan(s, tig + 4)][nt*8 + g].   TRANSPOSED = the backward product
// A W^T: element (k = co-side channel, n = ci) = W_term[n][k].
template <int CK, int CN, int TS, bool TRANSPOSED>
__device__ __forceinline__ void stage_weight_fragments(uint4* __restrict__ Bf, const float* __restrict__ W0, const float* __restrict__ W1,
                                                       const float* __restrict__ W2) {
    using G = SlabGeom<CK, TS>;
    constexpr int NT = CN / 8;
    for (int idx = threadIdx.x; idx < 3 * G::KS * NT * 32; idx += blockDim.x) {
        const int lane = idx & 31, rest = idx >> 5;
        const int nt = rest % NT, s = (rest / NT) % G::KS, term = rest / (NT * G::KS);
        const int g = lane >> 2, tig = lane & 3;
        const float* W = term == 0 ? W0 : (term == 1 ? W1 : W2);
        const int k0 = G::chan(s, tig), k1 = G::chan(s, tig + 4), n = nt * 8 + g;
        const float w0 = TRANSPOSED ? W[n * CK + k0] : W[k0 * CN + n];
        const float w1 = TRANSPOSED ? W[n * CK + k1] : W[k1 * CN + n];
        uint4 f;
        f.x = to_tf32(w0);
        f.y = to_tf32(w1);
        f.z = to_tf32(w0 - __uint_as_float(f.x));
        f.w = to_tf32(w1 - __uint_as_float(f.y));
        Bf[idx] = f;
    }
}

// Gather of one slab: acc[term][slot] (4 floats as two packed pairs), term 0 = own row, 1 = S0 row sum, 2 = S1 row sum.
// ment[p] = {internal column, (c1 << 16) | (c0 & 0xffff)} with the two integer coefficients as int16, columns ascending.
// Loads are unpredicated: a lane whose trajectory lies beyond b (ragged last slab) reads trajectory b-1 instead and its
// rows are never stored.
template <int C, int TS>
__device__ __forceinline__ void slab_accumulate(u64 (&acc)[3][SlabGeom<C, TS>::NL][2], int x, int pk,
                                                const u64 (&v)[SlabGeom<C, TS>::LPE][2]) {
    using G = SlabGeom<C, TS>;
    const float c0 = (float)(short)(pk & 0xffff), c1 = (float)(pk >> 16);
    const u64 q0 = bcast2(c0);
#pragma unroll
    for (int j = 0; j < G::LPE; ++j) {
        const int slot = x * G::LPE + j;
        ffma2(acc[1][slot][0], q0, v[j][0]);
        ffma2(acc[1][slot][1], q0, v[j][1]);
    }
    if (c1 != 0.f) {                                     // warp-uniform: most merged entries carry only the S0 coefficient
        const u64 q1 = bcast2(c1);
#pragma unroll
        for (int j = 0; j < G::LPE; ++j) {
            const int slot = x * G::LPE + j;
            ffma2(acc[2][slot][0], q1, v[j][0]);
            ffma2(acc[2][slot][1], q1, v[j][1]);
        }
    }
}

template <int C, int TS, int DEPTH = 2>
__device__ __forceinline__ void slab_gather(const float* __restrict__ H, unsigned rowbytes, const int32_t* __restrict__ mptr,
                                            const int2* __restrict__ ment, int E, int b, int e0, int t0,
                                            u64 (&acc)[3][SlabGeom<C, TS>::NL][2]) {
    using G = SlabGeom<C, TS>;
    const int lane = threadIdx.x & 31;
    const int tr = (4 * lane) / C;                       // trajectory of this lane inside a load block (same for lane ^ 4)
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < G::NL; ++i) acc[k][i][0] = acc[k][i][1] = 0ull;
    const char* Hb = reinterpret_cast<const char*>(H);
#pragma unroll
    for (int x = 0; x < G::EPS; ++x) {
        const int e = e0 + x;
        if (e >= E) break;                               // warp-uniform
        // byte offset of this lane inside an edge row, per slot of this edge (full slabs: off[j] = off[0] + 512 j)
        const char* P[G::LPE];
#pragma unroll
        for (int j = 0; j < G::LPE; ++j) {
            const int slot = x * G::LPE + j;
            const int lpos = (C == 32 && slot >= 2) ? (lane ^ 4) : lane;
            int t = t0 + j * G::Q + tr;
            t = t < b ? t : b - 1;
            P[j] = Hb + (size_t)((unsigned)(t * C + (4 * lpos) % C) * 4u);
            asm volatile("" : "+l"(P[j]));               // keep base + lane offset folded: one IMAD.WIDE per load address
        }
        int p = __ldg(mptr + e);
        const int p1 = __ldg(mptr + e + 1);
        {                                                // own row (term 0)
            const size_t r0 = (size_t)(unsigned)e * rowbytes;
#pragma unroll
            for (int j = 0; j < G::LPE; ++j)
                ldg128(reinterpret_cast<const float*>(P[j] + r0), acc[0][x * G::LPE + j][0], acc[0][x * G::LPE + j][1]);
        }
        if (DEPTH > 2)
            for (; p + DEPTH <= p1; p += DEPTH) {        // DEPTH neighbour rows in flight (kernels with registers to spare)
                int2 nd[DEPTH > 2 ? DEPTH : 1];
                u64 vd[DEPTH > 2 ? DEPTH : 1][G::LPE][2];
#pragma unroll
                for (int u = 0; u < DEPTH; ++u) nd[u] = __ldg(ment + p + u);
#pragma unroll
                for (int u = 0; u < DEPTH; ++u)
#pragma unroll
                    for (int j = 0; j < G::LPE; ++j)
                        ldg128(reinterpret_cast<const float*>(P[j] + (size_t)(unsigned)nd[u].x * rowbytes), vd[u][j][0], vd[u][j][1]);
#pragma unroll
                for (int u = 0; u < DEPTH; ++u) slab_accumulate<C, TS>(acc, x, nd[u].y, vd[u]);
            }
        for (; p + 2 <= p1; p += 2) {                    // two neighbour rows in flight
            const int2 na = __ldg(ment + p), nb = __ldg(ment + p + 1);
            u64 va[G::LPE][2], vb[G::LPE][2];
#pragma unroll
            for (int j = 0; j < G::LPE; ++j)
                ldg128(reinterpret_cast<const float*>(P[j] + (size_t)(unsigned)na.x * rowbytes), va[j][0], va[j][1]);
#pragma unroll
            for (int j = 0; j < G::LPE; ++j)
                ldg128(reinterpret_cast<const float*>(P[j] + (size_t)(unsigned)nb.x * rowbytes), vb[j][0], vb[j][1]);
            slab_accumulate<C, TS>(acc, x, na.y, va);
            slab_accumulate<C, TS>(acc, x, nb.y, vb);
        }
        if (p < p1) {
            const int2 na = __ldg(ment + p);
            u64 va[G::LPE][2];
#pragma unroll
            for (int j = 0; j < G::LPE; ++j)
                ldg128(reinterpret_cast<const float*>(P[j] + (size_t)(unsigned)na.x * rowbytes), va[j][0], va[j][1]);
            slab_accumulate<C, TS>(acc, x, na.y, va);
        }
    }
}

// A fragments of one term for all k-steps: fr[s][4] = {a0, a1, a2, a3} of k-step s (see SlabGeom).
template <int C, int TS>
__device__ __forceinline__ void slab_fragments(const u64 (&acc)[SlabGeom<C, TS>::NL][2], float (&fr)[SlabGeom<C, TS>::KS][4]) {
    using G = SlabGeom<C, TS>;
    float v[G::NL][4];
#pragma unroll
    for (int i = 0; i < G::NL; ++i) {
        unpack2(acc[i][0], v[i][0], v[i][1]);
        unpack2(acc[i][1], v[i][2], v[i][3]);
    }
    if (C == 32) {
        const bool odd = (threadIdx.x >> 2) & 1;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float send = odd ? v[q][j] : v[2 + q][j];
                const float keep = odd ? v[2 + q][j] : v[q][j];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
                fr[j][q] = keep;                         // a0 (q = 0: fragment row g) / a1 (q = 1: row g + 8), k = tig
                fr[j][2 + q] = recv;                     // a2 / a3, k = tig + 4
            }
    } else {
#pragma unroll
        for (int s = 0; s < G::KS; ++s) {
            fr[s][0] = v[0][2 * s];
            fr[s][1] = v[1][2 * s];
            fr[s][2] = v[0][2 * s + 1];
            fr[s][3] = v[1][2 * s + 1];
        }
    }
}

// (edge, trajectory) of this lane's fragment row r (0: row g, 1: row g + 8)
template <int C, int TS>
__device__ __forceinline__ void slab_row(int r, int e0, int t0, int& e, int& t) {
    using G = SlabGeom<C, TS>;
    const int g = (threadIdx.x & 31) >> 2;
    const int slot = C == 32 ? 2 * (g & 1) + r : r;
    const int trj = C == 32 ? (g >> 1) : g;
    e = e0 + slot / G::LPE;
    t = t0 + (slot % G::LPE) * G::Q + trj;
}

// d[nt] += A_term * W_term over all k-steps of one term (3xTF32)
template <int KS, int NT>
__device__ __forceinline__ void slab_mma_term(float (&d)[NT][4], const float (&fr)[KS][4], const uint4* __restrict__ Bterm) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        uint32_t ahi[4], alo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) split_tf32(fr[s][q], ahi[q], alo[q]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const uint4 w = Bterm[(s * NT + nt) * 32 + lane];
            mma_tf32(d[nt], alo, w.x, w.y);
            mma_tf32(d[nt], ahi, w.z, w.w);
            mma_tf32(d[nt], ahi, w.x, w.y);
        }
    }
}

