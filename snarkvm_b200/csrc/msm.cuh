// Pippenger multi-scalar multiplication over BLS12-377 G1 for sm_100a.
//
// Replaces the hot loops of the reference's VariableBase::msm
// (algorithms/src/msm/variable_base/mod.rs:30-49 → batched.rs:366-415):
//   batched_window  (batched.rs:328-364)  →  k_digits<0/1>                      (signed digits, counting sort by bucket)
//   batch_add       (batched.rs:175-325)  →  k_pair_level ×(0/2/4)              (Montgomery-trick affine pair levels)
//                                            k_bucket_accumulate(_dense)         (XYZZ sums of what is left, per work item)
//                                            k_partial_group_sum                 (hot buckets: fold item partials 32:1)
//   running sum     (batched.rs:356-361)  →  k_bucket_reduce / k_group_sum
//   window combine  (batched.rs:404-413)  →  host Horner over ≤ 24 window sums (host_ec.hpp)
//
// Design differences (the result is a group element, so any window size / digit
// recoding / coordinate system yields the same to_affine() image):
//   * signed c-bit digits  → 2^(c-1) buckets per window instead of 2^c − 1
//   * counting sort by (window, bucket) with global atomics instead of sort_unstable
//   * batched affine additions only for the first levels (where they fill the GPU), XYZZ mixed additions after
//   * every bucket is cut into work items of ≤ cap points and every pair level is split by OUTPUT elements, so a hot
//     bucket (all scalars equal, repeated bases — benches/msm/variable_base.rs:29-32 — or the carry-only top window)
//     cannot serialise a thread.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

struct MsmPlan {
    int c;              // window bits
    int nwin;           // number of windows = 253 / c + 1 (room for the signed-digit carry)
    uint32_t nbuckets;  // 2^(c-1) buckets per window (bucket value 1 .. 2^(c-1))
    uint32_t cap;       // max points per work item
    int levels;         // batched-affine pair levels run before the XYZZ accumulation (0 = gather + XYZZ only)
};

MsmPlan msm_make_plan(size_t npoints);

// Computes the per-window sums Σ_b b·S_{w,b} as XYZZ points (48 words each) into
// d_window_sums[plan.nwin][48].  All pointers are device pointers.  Scratch is taken
// from the stream-ordered pool of the current device.  Returns cudaError_t as int.
int msm_window_sums_device(uint32_t* d_window_sums, const MsmPlan& plan, const void* d_points, size_t stride,
                           const void* d_scalars, size_t npoints, cudaStream_t stream);

// Resident bases with precomputed tables 2^{c·w}·P_i (w < nwin; 128 B per record): all windows then share ONE bucket set, so
// wide windows are cheap (c = 22, 12 windows at 2^24 instead of 17 / 15).  Table size: npoints · nwin · 128 B.
MsmPlan msm_make_plan_precomputed(size_t npoints);
int msm_precompute_tables_device(uint32_t* d_table, const MsmPlan& plan, const void* d_points, size_t stride, size_t npoints, cudaStream_t stream);
// d_sum: ONE XYZZ point (192 B) = Σ scalars[i]·P_i over the first nscalars (≤ table_n) points
int msm_precomputed_sum_device(uint32_t* d_sum, const MsmPlan& plan, const uint32_t* d_table, size_t table_n, const void* d_scalars, size_t nscalars, cudaStream_t stream);

// Deterministic test/bench input: P_i = h(seed, i)·G with a 64-bit multiplier h
// (every point is in the prime-order subgroup because G is).  Writes the reference
// affine layout with the given stride.
int msm_generate_bases_device(void* d_points, size_t npoints, size_t stride, uint64_t seed, cudaStream_t stream);

// `.usrs` uncompressed points (96 B canonical each) → reference Affine images in HBM; *d_invalid counts points that are
// off the curve, out of range or badly flagged.
int srs_decode_device(void* d_out, size_t stride, const void* d_in, size_t npoints, uint32_t* d_invalid, cudaStream_t stream);

// out[i] = Σ_r in[r][i]  over `nranks` arrays of `count` XYZZ points (multi-GPU combine).
int xyzz_sum_ranks_device(uint32_t* d_out, const uint32_t* d_in, int nranks, int count, cudaStream_t stream);

// Once per device: keep freed scratch inside the stream-ordered pool (release threshold = max) so that
// steady-state calls never go back to the driver for their ~GB of workspace.
void ensure_pool_configured();

uint64_t launch_count();
void count_launch(int n = 1);

// Optional per-kernel CUDA-event timing on the launching stream (bench.py's roofline numbers).
enum ProfKind { PROF_MSM_SORT = 0, PROF_MSM_ACCUMULATE = 1, PROF_MSM_REDUCE = 2, PROF_NTT_PASS = 3, PROF_KINDS = 4 };
void prof_enable(bool on);
bool prof_enabled();
// records an event pair around whatever is enqueued on `stream` between begin and end
struct ProfScope {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t stream;
    int kind;
    ProfScope(int kind, cudaStream_t stream);
    ~ProfScope();
};
// waits for all recorded pairs of `kind`, returns their summed duration and count, and clears them
int prof_collect(int kind, double* total_ms, uint64_t* count);

}  // namespace b200
