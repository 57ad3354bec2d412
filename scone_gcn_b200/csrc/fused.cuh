// fused.cuh — state and entry points of the trajectory-fused pipeline (scone_fused.cu), used by scone_model.cu.
#pragma once
#include "common.cuh"

constexpr int kFusedMaxL = 3;             // conv layers the fused compute kernel keeps weight-gradient tiles in registers for
constexpr int kFusedHdrW = 16;            // ints per trajectory header
constexpr int kFusedFlagOverflow = 1;
constexpr int kFusedFlagRetry = 2;        // tier 0 of the plan kernel overflowed its tables: tier 1 redoes the trajectory

// Per-trajectory header written by fused_plan_kernel (word offsets into the program arena):
//   [0] flags   [1..3] live rows of layers 1..3   [4] layer-1 scalars {x, S0 x, S1 x} per row
//   [5],[6] forward programs of layers 2, 3: rowptr[n_l + 1] then int2 entries {row below | own << 31, (c1 << 16) | c0}
//   [7],[8] transposed programs of layers 2, 3 (rows = live rows of layer l - 1, entries = rows of layer l)
//   [9] readout: rowptr[D + 1] then int2 {row of H_L | slot << 16, sign bits}   [10] pairs   [11] hash entries (flows + cone)
//   [12] unused   [13] flow entries
struct FusedState {
    int L = 0, C = 0;
    int64_t n_params = 0;
    int bound_cone = 0, bound_list = 0;   // static bounds of the complex: |T_0| (hash entries) / |T_1| (rows of a layer) of any last node
    unsigned* d_cone_ptr = nullptr;       // cone table: T_0 of every node, entries edge | level << 30
    uint32_t* d_cone_ent = nullptr;
    unsigned long long cone_entries = 0;
    int HS = 0, LV = 0, EC = 0, hshift = 0;        // plan tables, tier 1: sized by the bounds
    int HS0 = 0, LV0 = 0, EC0 = 0, hshift0 = 0;   // tier 0: sized for the cones of 99 % of the nodes
    bool two_tiers = false;
    int flow_room = 0;
    size_t plan_smem0 = 0;
    unsigned long long worst_words = 0;
    int* d_retry = nullptr;
    size_t plan_smem = 0, traj_smem_small = 0, traj_smem_big = 0;
    int cap_rows = 0, big_rows = 0, grid_small = 0, grid_big = 0, chunk = 0;
    size_t scratch_stride = 0;
    unsigned long long arena_words = 0;
    int* d_hdr = nullptr;
    uint32_t* d_arena = nullptr;
    unsigned long long* d_bump = nullptr;
    float* d_partial = nullptr;
    float* d_scratch = nullptr;
    int* d_stats = nullptr;
    unsigned long long* d_rows_done = nullptr;
    // planned set (scone_fused_plan_set): headers + programs of a whole dataset, kept across calls
    int* d_set_hdr = nullptr;
    uint32_t* d_set_arena = nullptr;
    unsigned long long* d_set_bump = nullptr;
    unsigned long long set_arena_words = 0;
    int set_cap = 0, set_n = 0;
};

bool scone_fused_supported(const scone_complex* cx, int n_layers, const int32_t* hidden);
int scone_fused_create(const scone_complex* cx, int L, int C, int mb, int64_t n_params, FusedState** out);
void scone_fused_destroy(FusedState* f);
int scone_fused_run(const scone_complex* cx, FusedState* f, int act, int b, const int32_t* traj_ptr, const int32_t* flow_edge,
                    const float* flow_val, const int32_t* last_nodes, const float* W, const int64_t* w_off, float* logprobs,
                    const int32_t* target_idx, const float* mask, float* grad, int* overflow, bool count_rows, cudaStream_t st);
int scone_fused_plan_set(const scone_complex* cx, FusedState* f, int B, const int32_t* traj_ptr, const int32_t* flow_edge,
                         const float* flow_val, const int32_t* last_nodes, int* overflow, cudaStream_t st);
int scone_fused_run_planned(const scone_complex* cx, FusedState* f, int act, int n, const int32_t* rows_dev, const float* W, const int64_t* w_off,
                            float* logprobs, const int32_t* target_idx, const float* mask, float* grad, bool count_rows, cudaStream_t st);
int scone_fused_last_retries(FusedState* f);
int scone_fused_read(FusedState* f, int t, int* hdr_out, unsigned off, int words, uint32_t* arena_out);
