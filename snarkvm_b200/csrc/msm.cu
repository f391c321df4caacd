// See msm.cuh for the map from reference functions to kernels.
#include "msm.cuh"

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_scan.cuh>

#define FF_CALL_MUL 1
#include "ec.cuh"
#include "cta_inverse.cuh"
#include "quad.cuh"

namespace b200 {

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n); }

static std::atomic<bool> g_prof{false};
static std::mutex g_prof_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_recs[PROF_KINDS];
void prof_enable(bool on) { g_prof.store(on); }
bool prof_enabled() { return g_prof.load(); }
ProfScope::ProfScope(int kind_, cudaStream_t stream_) : stream(stream_), kind(kind_) {
    if (!g_prof.load()) return;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
    cudaEventRecord(a, stream);
}
ProfScope::~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, stream);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof_recs[kind].emplace_back(a, b);
}
int prof_collect(int kind, double* total_ms, uint64_t* count) {
    if (kind < 0 || kind >= PROF_KINDS) return (int)cudaErrorInvalidValue;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> recs;
    { std::lock_guard<std::mutex> lock(g_prof_mu); recs.swap(g_prof_recs[kind]); }
    double tot = 0; int rc = 0;
    for (auto& r : recs) {
        float ms = 0;
        cudaError_t e = cudaEventSynchronize(r.second);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.first, r.second);
        if (e != cudaSuccess) rc = (int)e; else tot += ms;
        cudaEventDestroy(r.first); cudaEventDestroy(r.second);
    }
    if (total_ms) *total_ms = tot;
    if (count) *count = recs.size();
    return rc;
}

#define CUDA_TRY(x)                                   \
    do {                                              \
        cudaError_t e_ = (x);                         \
        if (e_ != cudaSuccess) { rc = (int)e_; goto done; } \
    } while (0)

static int ceil_log2(size_t x) { int l = 0; size_t v = x > 1 ? x - 1 : 0; while (v) { l++; v >>= 1; } return l; }
// log2 rounded to the NEAREST integer (in the log domain): the plans below were swept at powers of two, and a size just above one
// — a 2^20-coefficient polynomial plus four blinding terms — belongs to that power's plan, not to the next one's
// (round 1 of the Varuna prover committed w with the 2^21 plan: 15.5 ms instead of 9.8).
static int plan_log2(size_t x) {
    int l = ceil_log2(x < 2 ? 2 : x);
    if (l > 1 && (double)x < 0.70710678118 * (double)((size_t)1 << l)) l--;
    return l;
}

MsmPlan msm_make_plan(size_t npoints) {
    MsmPlan p;
    int lg = plan_log2(npoints);
    // Window bits from a sweep on B200 (tools/tune_msm.py, profiles/tune_msm_r1.log): wider windows mean fewer
    // bucket additions (n·W) but more buckets to reduce and shorter, more divergent bucket runs.
    int c = lg <= 8 ? 4 : lg <= 12 ? lg - 4 : lg <= 18 ? 11 : lg == 19 ? 13 : lg == 20 ? 15 : lg <= 22 ? 16 : 17;
    if (const char* e = getenv("SNARKVM_B200_MSM_C")) { int v = atoi(e); if (v >= 2 && v <= 24) c = v; }
    p.c = c;
    p.nwin = 253 / c + 1;
    p.nbuckets = 1u << (c - 1);
    // Aim for ≥ ~300k work items so 148 SMs × (2 × 256-thread CTAs) see several waves, but keep
    // items long enough (≥ 16 points) that the per-item overhead stays in the noise.
    size_t total = npoints * (size_t)p.nwin;
    size_t cap = total / 300000 + 1;
    if (cap < 16) cap = 16;
    if (const char* e = getenv("SNARKVM_B200_MSM_CAP")) { long v = atol(e); if (v >= 1) cap = (size_t)v; }
    p.cap = (uint32_t)cap;
    // Batched-affine pair levels before the XYZZ accumulation pay off only when every level still fills the
    // GPU (tools/ab_pair.py sweep after the CTA-shared inversion): none below 2^20 points, 1 at 2^20, 4 from 2^21.
    // (re-swept in round 2 with the record-scatter sort, profiles/r2g_ab_records.log, r2h_ab_karatsuba_build.log: 5 levels from
    // 2^23, 4 at 2^21–2^22, 3 at 2^20, 2 at 2^19 with c = 13)
    int levels = lg >= 23 ? 5 : lg >= 21 ? 4 : lg == 20 ? 3 : lg == 19 ? 2 : 0;
    while (levels > 0 && ((npoints >> (c - 1)) >> levels) < 2) levels--;
    if (const char* e = getenv("SNARKVM_B200_MSM_LEVELS")) { int v = atoi(e); if (v >= 0 && v <= 16) levels = v; }
    p.levels = levels;
    return p;
}

// Plan for `njobs` sums sharing one base set in one pass: the window size follows the LARGEST job (bucket load), the
// number of pair levels follows the TOTAL work (a level is worth it while it fills the machine) as long as the buckets
// of the largest job still hold a few points after the halvings.
MsmPlan msm_make_plan_batch(size_t max_n, size_t total_n) {
    MsmPlan p = msm_make_plan(max_n);
    if (total_n > max_n) {
        int lgt = plan_log2(total_n);
        int levels = lgt >= 23 ? 5 : lgt >= 21 ? 4 : lgt == 20 ? 3 : lgt == 19 ? 2 : 0;
        // (a level the total asks for beyond the largest job's own plan must leave ≥ 4 points per bucket: 8 × 2^20 coefficients with
        // c = 15 hold 64 per bucket — 4 levels 57.7 ms, 5 levels 60.4 ms, profiles/r2ag_batch.log)
        while (levels > p.levels && ((max_n >> (p.c - 1)) >> levels) < 4) levels--;
        while (levels > 0 && ((max_n >> (p.c - 1)) >> levels) < 2) levels--;
        if (const char* e = getenv("SNARKVM_B200_MSM_LEVELS")) { int v = atoi(e); if (v >= 0 && v <= 16) levels = v; }
        if (levels > p.levels) p.levels = levels;
        size_t cap = total_n * (size_t)p.nwin / 300000 + 1;
        if (cap < 16) cap = 16;
        if (const char* e = getenv("SNARKVM_B200_MSM_CAP")) { long v = atol(e); if (v >= 1) cap = (size_t)v; }
        p.cap = (uint32_t)cap;
    }
    return p;
}

// ---------------------------------------------------------------------------
// Signed-digit recoding of a canonical 253-bit scalar (8 little-endian u32 words).
// digit_w ∈ [-2^(c-1), 2^(c-1)]; returns magnitude and sign for window w given the
// running carry (sequential over w).
// ---------------------------------------------------------------------------
// flat != 0 (precomputed tables 2^{c·w}·P_i): every window feeds the SAME bucket set and the entry names record
// w·flat + i of the table instead of point i.  slot_base: first counter of this job's bucket sets; index_base: position of
// this scalar vector's first point in the call's dense base array.  MONT: the scalars are Montgomery Fr (polynomial
// coefficients) and are converted here (to_bigint, kzg10/mod.rs:469-474).  Scalars with bits 253..255 set are outside
// what nwin windows cover: they raise bit 0 of *flags and the caller returns an error instead of a wrong point.
template <bool SCATTER, bool MONT>
__global__ void __launch_bounds__(256) k_digits(const uint32_t* __restrict__ scalars, size_t n, int c, int nwin,
                                                uint32_t nbuckets, uint32_t* __restrict__ counters /* hist or cursors */,
                                                uint32_t* __restrict__ sorted, size_t flat /* 0, or the table's points per window */,
                                                uint32_t slot_base, uint32_t index_base, uint32_t* __restrict__ flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    {
        const uint4* q = reinterpret_cast<const uint4*>(scalars + 8 * i);
        uint4 a = __ldg(q), b = __ldg(q + 1);
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    }
    if (MONT) {
        Fr x;
#pragma unroll
        for (int k = 0; k < 8; k++) x.v[k] = s[k];
        x = x.from_mont();
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = x.v[k];
    }
    if (!SCATTER && (s[7] >> 29)) atomicOr(flags, 1u);
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < nwin; w++) {
        int bit = w * c;
        // dynamic register-array indexing would spill: select the two words with a small switch-free scan
        int wi = bit >> 5, sh = bit & 31;
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { if (k == wi) lo = s[k]; if (k == wi + 1) hi = s[k]; }
        uint32_t raw = (__funnelshift_r(lo, hi, sh) & ((1u << c) - 1u)) + carry;
        uint32_t neg = raw > half ? 1u : 0u;
        uint32_t mag = neg ? (1u << c) - raw : raw;
        carry = neg;
        if (mag != 0u) {
            uint32_t slot = slot_base + (flat ? 0u : (uint32_t)w * nbuckets) + (mag - 1u);
            if (SCATTER) {
                uint32_t pos = atomicAdd(&counters[slot], 1u);
                sorted[pos] = (uint32_t)(flat ? (size_t)w * flat + i : (size_t)index_base + i) | (neg << 31);
            } else {
                atomicAdd(&counters[slot], 1u);
            }
        }
    }
}

__global__ void k_items_per_bucket(const uint32_t* __restrict__ hist, uint32_t* __restrict__ items, uint32_t total_buckets, uint32_t cap,
                                   uint32_t* __restrict__ hot /* may be null; hot[0] = count, hot[1 …] = buckets with more than `keep` items */, uint32_t hot_max,
                                   uint32_t keep) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_buckets) {
        const uint32_t it = (hist[i] + cap - 1u) / cap;
        items[i] = it;
        if (hot != nullptr && it > keep) { const uint32_t pos = atomicAdd(hot, 1u); if (pos < hot_max) hot[1u + pos] = i; }
    } else if (i == total_buckets) items[i] = 0;
}

// Dense points are 96 B (x, y Montgomery); infinity is encoded as (0, 0), which is not on y² = x³ + 1.
static constexpr int DENSE_WORDS = 24;
static constexpr int BASE_WORDS = 32;       // level-0 copy of the bases: 96 B padded to one 128-byte line per point

struct DensePoint { Fq x, y; bool inf; };
FF_DEV DensePoint load_dense(const uint32_t* p) {
    DensePoint d; d.x = Fq::load_ldg(p); d.y = Fq::load_ldg(p + 12);
    d.inf = d.x.is_zero() && d.y.is_zero();
    return d;
}
FF_DEV void store_dense(uint32_t* p, const DensePoint& d) {
    if (d.inf) { Fq z = Fq::zero(); z.store(p); z.store(p + 12); }
    else { d.x.store(p); d.y.store(p + 12); }
}
// Sort pass of the pair-level path (round 2): instead of an index array that level 0 would have to chase through a
// random gather (ncu, profiles/r2f_*: the gathering level ran the multiplier at 57 % where the dense levels reach 83 % —
// 32 DRAM lines per LDGSTS instruction, MIO and scoreboard stalls with only 4 warps per scheduler to hide them), the
// scatter writes the RECORDS themselves: point i is read once, and for every window its 96-byte (x, ±y) image goes to the
// slot the bucket cursor hands out.  Level 0 then reads a dense, already signed array exactly like the levels above it.
// The CTA stages its 256 points (x, y, −y) in shared memory and writes each window's records cooperatively, six
// consecutive lanes per record, so a store instruction touches 6 lines instead of 32.
// Windows [w_lo, w_hi) of this segment belong to the current group of bucket sets; positions are relative to *pos_base.
static constexpr uint32_t REC_NONE = 0xffffffffu;
// FLAT (precomputed tables): window w of point i takes record w·table_n + i of the table (staged per window) and every window
// feeds the job's single bucket set.
template <bool MONT, bool FLAT>
__global__ void __launch_bounds__(256) k_scatter_records(const uint32_t* __restrict__ scalars, size_t n, const uint8_t* __restrict__ points,
                                                         size_t stride, const uint32_t* __restrict__ table, size_t table_n, int c, uint32_t nbuckets,
                                                         uint32_t* __restrict__ cursors, uint32_t slot_base, int w_lo, int w_hi,
                                                         const uint32_t* __restrict__ pos_base_ptr, uint4* __restrict__ dense0) {
    __shared__ uint4 sh_rec[256 * 9];                       // per point: x (3 × 16 B), y (3), −y (3)
    __shared__ uint32_t sh_pos[2][256];
    const uint32_t tid = threadIdx.x;
    const size_t i = (size_t)blockIdx.x * 256 + tid;
    const bool live = i < n;
    const uint32_t pos_base = __ldg(pos_base_ptr);
    uint32_t s[8];
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = 0u;
    auto stage = [&](Fq x, Fq y, bool inf) {
        Fq yn = y.neg();
        if (inf) { x = Fq::zero(); y = Fq::zero(); yn = Fq::zero(); }      // (0, 0) is the dense encoding of infinity
        uint4* r = sh_rec + tid * 9;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            r[k] = make_uint4(x.v[4 * k], x.v[4 * k + 1], x.v[4 * k + 2], x.v[4 * k + 3]);
            r[3 + k] = make_uint4(y.v[4 * k], y.v[4 * k + 1], y.v[4 * k + 2], y.v[4 * k + 3]);
            r[6 + k] = make_uint4(yn.v[4 * k], yn.v[4 * k + 1], yn.v[4 * k + 2], yn.v[4 * k + 3]);
        }
    };
    if (live) {
        const uint4* q = reinterpret_cast<const uint4*>(scalars + 8 * i);
        uint4 a = __ldg(q), b = __ldg(q + 1);
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
        if (MONT) {
            Fr x;
#pragma unroll
            for (int k = 0; k < 8; k++) x.v[k] = s[k];
            x = x.from_mont();
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = x.v[k];
        }
        if (!FLAT) {
            AffinePoint pt = load_affine(points, stride, i);
            stage(pt.x, pt.y, pt.inf);
        }
    }
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0;
    // One window at a time (issuing the cursor atomics of 8 windows back to back before one barrier was measured slower:
    // 10.5 vs 9.0 ms at 2^24, profiles/r2h_ab_karatsuba_build.log — co-resident CTAs in different phases already overlap).
    for (int w = 0; w < w_hi; w++) {
        const int bit = w * c, wi = bit >> 5, sh = bit & 31;
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { if (k == wi) lo = s[k]; if (k == wi + 1) hi = s[k]; }
        const uint32_t raw = (__funnelshift_r(lo, hi, sh) & ((1u << c) - 1u)) + carry;
        const uint32_t neg = raw > half ? 1u : 0u;
        const uint32_t mag = neg ? (1u << c) - raw : raw;
        carry = neg;
        if (w < w_lo) continue;
        if (FLAT) {
            if (w > w_lo) __syncthreads();                  // the previous window's records have been written out
            if (live && mag != 0u) {
                const DensePoint d = load_dense(table + ((size_t)w * table_n + i) * BASE_WORDS);
                stage(d.x, d.y, d.inf);
            }
        }
        uint32_t pos = REC_NONE;
        if (live && mag != 0u) pos = (atomicAdd(&cursors[slot_base + (FLAT ? 0u : (uint32_t)w * nbuckets) + (mag - 1u)], 1u) - pos_base) | (neg << 31);
        uint32_t* my_pos = sh_pos[w & 1];
        my_pos[tid] = pos;
        __syncthreads();                                    // also orders the staging of sh_rec before the reads
        for (uint32_t k = tid; k < 256u * 6u; k += 256u) {
            const uint32_t r = k / 6u, part = k - 6u * r;
            const uint32_t pp = my_pos[r];
            if (pp == REC_NONE) continue;
            const uint32_t src = part < 3u ? part : ((pp >> 31) ? 3u : 0u) + part;       // x0..2 | y0..2 or (−y)0..2
            dense0[(size_t)(pp & 0x7fffffffu) * 6u + part] = sh_rec[r * 9u + src];
        }
        // the next window writes the other half of sh_pos; its barrier orders this window's reads before the window after
    }
}

// A dense base record: x, y (Montgomery) in one 128-byte line, infinity encoded as (0, 0) — written by k_densify_bases
// (per call) or k_precompute_tables (per SRS).
FF_DEV AffinePoint load_record(const uint32_t* __restrict__ records, uint32_t idx) {
    const uint32_t* p = records + (size_t)idx * 32;
    AffinePoint a;
    a.x = Fq::load_ldg(p); a.y = Fq::load_ldg(p + 12);
    a.inf = a.x.is_zero() && a.y.is_zero();
    return a;
}
// One thread per work item (a run of ≤ cap sorted entries of one bucket): XYZZ mixed additions
// of gathered affine bases.  partial[item] receives the item's sum.
// 128-thread CTAs, 4 per SM (≤ 128 registers/thread, a few hundred bytes of spill): measured on B200 at
// 2^24 points — 129 ms vs 140 ms for 256×1 at 207 registers (profiles/README.md).  The kernel is bound by
// the IMAD pipe, and 16 warps/SM hide its latency better than 8.
#ifndef MSM_ACC_THREADS
#define MSM_ACC_THREADS 128
#define MSM_ACC_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(MSM_ACC_THREADS, MSM_ACC_MINBLOCKS) k_bucket_accumulate(const uint32_t* __restrict__ records /* 128-byte dense bases or table */,
                                                            const uint32_t* __restrict__ sorted,
                                                            const uint32_t* __restrict__ bucket_start /* [TB+1] */,
                                                            const uint32_t* __restrict__ item_start /* [TB+1] */,
                                                            uint32_t total_buckets, uint32_t cap, uint32_t* __restrict__ partial) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t total_items = item_start[total_buckets];
    if (t >= total_items) return;
    // upper_bound(item_start, t) - 1 : the bucket whose item range contains t
    uint32_t lo = 0, hi = total_buckets;          // invariant: item_start[lo] <= t < item_start[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (item_start[mid] <= t) lo = mid; else hi = mid;
    }
    uint32_t wb = lo;
    uint32_t seg = t - item_start[wb];
    uint32_t b0 = bucket_start[wb], b1 = bucket_start[wb + 1];
    uint32_t s0 = b0 + seg * cap;
    uint32_t s1 = s0 + cap < b1 ? s0 + cap : b1;

    XYZZ acc = XYZZ::infinity();
    // software pipeline: fetch entry k+1 while adding entry k
    uint32_t e = sorted[s0];
    AffinePoint p = load_record(records, e & 0x7fffffffu);
    for (uint32_t k = s0; k < s1; k++) {
        uint32_t e_cur = e;
        AffinePoint p_cur = p;
        if (k + 1 < s1) { e = sorted[k + 1]; p = load_record(records, e & 0x7fffffffu); }
        acc.add_affine(p_cur, (e_cur >> 31) != 0u);
    }
    acc.store(partial + (size_t)t * XYZZ_WORDS);
}

// =================================================================================================
// Batched-affine pair levels (the reference's own idea — batch_add, batched.rs:175-325: pair up the
// points of a bucket, add all pairs with ONE field inversion per batch via Montgomery's trick,
// affine.rs:224-273 — restated for the GPU).  One level halves every bucket: output element i of a
// bucket is in[2i] + in[2i+1] (or a copy of in[2i] when the count is odd).  A thread owns T consecutive
// OUTPUT elements (across bucket boundaries, so hot buckets are spread over many threads), walks them
// forward accumulating the running product of the denominators (x2 − x1, or 2·y1 for a doubling) into
// `prefix`, inverts once (Fermat, ≈ 570 Fq mul amortised over T = 64…1024 additions), then walks them
// backward peeling off one inverse per pair: 6 Fq mul per addition instead of 10 for an XYZZ mixed add.
// Dense points are 96 B (x, y Montgomery); infinity is encoded as (0, 0), which is not on y² = x³ + 1.
// =================================================================================================
// Level-0 inputs are gathered through `sorted` from a dense, 128-byte-aligned copy of the bases made once
// per call by k_densify_bases: a gathered point then costs one DRAM line instead of the two or three that
// the reference's 104-byte stride straddles.
template <bool GATHER>
FF_DEV DensePoint load_level_input(const uint32_t* __restrict__ dense_bases, const uint32_t* __restrict__ sorted,
                                   const uint32_t* __restrict__ dense_in, uint32_t idx) {
    DensePoint d;
    if (GATHER) {
        uint32_t e = sorted[idx];
        d = load_dense(dense_bases + (size_t)(e & 0x7fffffffu) * BASE_WORDS);
        if ((e >> 31) && !d.inf) d.y = d.y.neg();
    } else {
        d = load_dense(dense_in + (size_t)idx * DENSE_WORDS);
    }
    return d;
}
// address of the record (forward pass reads x only: the denominator x2 − x1 needs nothing else unless the x's collide)
template <bool GATHER>
FF_DEV const uint32_t* level_input_ptr(const uint32_t* __restrict__ dense_bases, const uint32_t* __restrict__ sorted,
                                       const uint32_t* __restrict__ dense_in, uint32_t idx) {
    if (GATHER) return dense_bases + (size_t)(sorted[idx] & 0x7fffffffu) * BASE_WORDS;
    return dense_in + (size_t)idx * DENSE_WORDS;
}
__global__ void k_densify_bases(const uint8_t* __restrict__ points, size_t stride, size_t n, uint32_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    AffinePoint a = load_affine(points, stride, i);
    DensePoint d; d.x = a.x; d.y = a.y; d.inf = a.inf;
    store_dense(out + i * BASE_WORDS, d);
}
// Precomputed tables for resident bases: record (w, i) = 2^{c·w}·P_i as a dense 128-byte affine record, w < nwin.
// One thread per point walks the doubling chain in XYZZ; every multiple is normalised with ONE field inversion shared by
// the 128 threads of the CTA (cta_inverse.cuh) instead of a Fermat ladder per record: ≈ 9·c + 12 Fq mul per record instead of
// 9·c + 580 (round 1: 6.3 s for the 2^24-point tables).
__global__ void __launch_bounds__(CTA_INV_THREADS) k_precompute_tables(const uint8_t* __restrict__ points, size_t stride, size_t n, int c, int nwin,
                                                            uint32_t* __restrict__ table) {
    __shared__ uint4 sh_inv[CTA_INV_SMEM_BYTES / 16];
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    AffinePoint a;
    a.x = Fq::zero(); a.y = Fq::zero(); a.inf = true;
    if (live) a = load_affine(points, stride, i);
    DensePoint d; d.x = a.x; d.y = a.y; d.inf = a.inf;
    if (live) store_dense(table + i * BASE_WORDS, d);
    XYZZ q = XYZZ::from_affine(a);
    for (int w = 1; w < nwin; w++) {
        for (int k = 0; k < c; k++) q.dbl();
        const bool inf = q.is_inf();
        Fq z = inf ? Fq::one() : q.ZZ * q.ZZZ;
        Fq iz = cta_shared_inverse_by(z, reinterpret_cast<uint32_t*>(sh_inv), w & 3);
        __syncthreads();                                    // the next round overwrites the shared area
        if (inf) { d.x = Fq::zero(); d.y = Fq::zero(); d.inf = true; }
        else { d.x = q.X * (iz * q.ZZZ); d.y = q.Y * (iz * q.ZZ); d.inf = false; }
        if (live) store_dense(table + ((size_t)w * n + i) * BASE_WORDS, d);
    }
}

static constexpr int PAIR_THREADS = CTA_INV_THREADS;
enum PairKind { PAIR_COPY1 = 0, PAIR_COPY2 = 1, PAIR_INF = 2, PAIR_ADD = 3, PAIR_DBL = 4 };
FF_DEV int classify_pair(const DensePoint& P, const DensePoint& Q, bool has2, Fq& d) {
    if (!has2 || Q.inf) return PAIR_COPY1;
    if (P.inf) return PAIR_COPY2;
    if (P.x == Q.x) {
        if (P.y == Q.y && !P.y.is_zero()) { d = P.y.dbl(); return PAIR_DBL; }
        return PAIR_INF;                                   // P + (−P)  (or a 2-torsion point doubled)
    }
    d = Q.x - P.x;
    return PAIR_ADD;
}

template <bool GATHER>
__global__ void __launch_bounds__(PAIR_THREADS, 4) k_pair_level(const uint32_t* __restrict__ dense_bases,
                                                        const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ dense_in,
                                                        const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out,
                                                        uint32_t total_buckets, uint32_t T, uint32_t* __restrict__ prefix,
                                                        uint32_t* __restrict__ dense_out) {
    __shared__ uint4 sh_inv4[PAIR_THREADS * 3];           // one Fq per thread for the CTA-wide shared inversion
    uint32_t* sh_inv = reinterpret_cast<uint32_t*>(sh_inv4);
    const uint32_t total = off_out[total_buckets];
    const uint64_t o0_64 = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * T;
    // threads past the end stay for the barriers of the shared inversion with an empty range
    const uint32_t o0 = o0_64 < total ? (uint32_t)o0_64 : total;
    const uint32_t o1 = (o0_64 + T < total) ? o0 + T : total;
    uint32_t lo = 0, hi = total_buckets;                  // off_out[lo] <= o0 < off_out[hi]
    if (o0 < o1) while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (off_out[mid] <= o0) lo = mid; else hi = mid; }
    uint32_t b = lo;

    // ---- forward: running product of denominators (x coordinates only on the common path) ----
    Fq run = Fq::one();
    for (uint32_t o = o0; o < o1; o++) {
        while (o >= off_out[b + 1]) b++;
        const uint32_t i = o - off_out[b], base_in = off_in[b], cnt = off_in[b + 1] - base_in;
        const bool has2 = 2 * i + 1 < cnt;
        if (has2) {
            const uint32_t* pp = level_input_ptr<GATHER>(dense_bases, sorted, dense_in, base_in + 2 * i);
            const uint32_t* qp = level_input_ptr<GATHER>(dense_bases, sorted, dense_in, base_in + 2 * i + 1);
            Fq x1 = Fq::load_ldg(pp), x2 = Fq::load_ldg(qp);
            if (x1 == x2 || x1.is_zero() || x2.is_zero()) {
                // rare: equal x (doubling / cancellation) or a possible infinity — take the full path
                DensePoint P = load_level_input<GATHER>(dense_bases, sorted, dense_in, base_in + 2 * i);
                DensePoint Q = load_level_input<GATHER>(dense_bases, sorted, dense_in, base_in + 2 * i + 1);
                Fq d;
                int kind = classify_pair(P, Q, true, d);
                if (kind >= PAIR_ADD) run = run * d;
            } else {
                run = run * (x2 - x1);
            }
        }
        run.store(prefix + (size_t)o * 12);
    }
    Fq inv = cta_shared_inverse(run, sh_inv);
    // ---- backward: one inverse per pair, then the affine addition / doubling ----
    for (uint32_t o = o1; o-- > o0;) {
        while (o < off_out[b]) b--;
        const uint32_t i = o - off_out[b], base_in = off_in[b], cnt = off_in[b + 1] - base_in;
        const bool has2 = 2 * i + 1 < cnt;
        DensePoint P = load_level_input<GATHER>(dense_bases, sorted, dense_in, base_in + 2 * i);
        DensePoint Q = P;
        if (has2) Q = load_level_input<GATHER>(dense_bases, sorted, dense_in, base_in + 2 * i + 1);
        Fq d;
        int kind = classify_pair(P, Q, has2, d);
        DensePoint R;
        if (kind == PAIR_COPY1) R = P;
        else if (kind == PAIR_COPY2) R = Q;
        else if (kind == PAIR_INF) { R.inf = true; R.x = Fq::zero(); R.y = Fq::zero(); }
        else {
            Fq inv_d = (o == o0) ? inv : inv * Fq::load(prefix + (size_t)(o - 1) * 12);
            inv = inv * d;
            Fq lambda;
            if (kind == PAIR_ADD) lambda = (Q.y - P.y) * inv_d;
            else { Fq xx = P.x.sqr(); lambda = (xx.dbl() + xx) * inv_d; }
            Fq x3 = lambda.sqr() - P.x - Q.x;                // Q.x == P.x in the doubling case
            R.y = lambda * (P.x - x3) - P.y;
            R.x = x3;
            R.inf = false;
        }
        store_dense(dense_out + (size_t)o * DENSE_WORDS, R);
    }
}


// =================================================================================================
// Pair level, second design (round 2): warp-interleaved outputs + a shared-memory operand ring.
//
// v1 above gives every THREAD a run of T consecutive outputs: the 32 lanes of a load instruction then touch 32
// different DRAM pages / L1 lines (sorted[], prefix[], the level's dense input), every record is fetched with the
// multiplier idle behind a 4-deep dependent chain (off_out → off_in → sorted → record), and ncu shows 2.5–4.3
// long-scoreboard stall cycles per issued instruction at 65–72 % of the multiplier pipe.  Here a WARP owns 32·T
// consecutive outputs and lane l takes outputs W0 + 32·j + l, j < T — any partition works for Montgomery's trick —
// so sorted[], prefix[] and the dense inputs/outputs of a step are contiguous across the warp, and the operands of
// step j+1 are already on their way into shared memory (cp.async / LDGSTS, 16-byte granules into a per-warp
// [stage][chunk][lane] ring, conflict-free for LDS.128) while step j multiplies: the index chain runs two steps
// ahead (bucket walk + sorted[] entry), the record copies one step ahead, and the multiplier calls read their
// operands from shared memory when they need them, so nothing but the running inverse is live across a call.
// Each record still crosses DRAM once per pass (forward: x only; backward: x and y); prefix products go to HBM
// coalesced (48 B per output) and come back one step late behind the first multiplication of the step.
// =================================================================================================
FF_DEV uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FF_DEV void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
FF_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
FF_DEV void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
FF_DEV void cp_async_wait_0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
FF_DEV void cp_async_wait_3() { asm volatile("cp.async.wait_group 3;" ::: "memory"); }

static constexpr int RING_CHUNKS = 12;                       // x1 y1 x2 y2, three 16-byte chunks each
static constexpr int RING_STAGE_U4 = RING_CHUNKS * 32;       // uint4 per warp per stage (6 KiB)
static constexpr int PAIR2_SMEM = PAIR_THREADS * 48 + 4 * 2 * RING_STAGE_U4 * 16;   // shared inversion + 4 warps × 2 stages

FF_DEV Fq ring_fq(const uint4* slot) {                       // slot = &ring[(stage·12 + first chunk)·32 + lane]
    Fq r;
#pragma unroll
    for (int i = 0; i < 3; i++) { uint4 t = slot[i * 32]; r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w; }
    return r;
}

// Output descriptors, written once per level by k_pair_desc (one thread per output, warp-coherent binary search for the
// bucket) so that the pair kernel never chases off_out → off_in → sorted in its instruction stream: it reads 8 bytes per
// output, coalesced, two steps ahead.
//   level 0:  p, q = the two sorted entries (point index | sign << 31); q = NONE when the output has a single input
//   above:    p = index of the first input in dense_in; q = NONE / anything else
static constexpr uint32_t PAIR_NONE = 0xffffffffu;
template <bool GATHER>
__global__ void __launch_bounds__(256) k_pair_desc(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ off_in,
                                                   const uint32_t* __restrict__ off_out, uint32_t total_buckets, uint2* __restrict__ desc,
                                                   const uint32_t* __restrict__ in_base_ptr /* dense inputs: *in_base_ptr = position of element 0, or null */) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= __ldg(off_out + total_buckets)) return;
    uint32_t lo = 0, hi = total_buckets;                  // off_out[lo] <= o < off_out[hi]
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (__ldg(off_out + mid) <= o) lo = mid; else hi = mid; }
    const uint32_t i = o - __ldg(off_out + lo), base_in = __ldg(off_in + lo), cnt = __ldg(off_in + lo + 1) - base_in;
    const uint32_t idx = base_in + 2u * i;
    const bool has2 = 2u * i + 1u < cnt;
    uint2 d;
    if (GATHER) { d.x = __ldg(sorted + idx); d.y = has2 ? __ldg(sorted + idx + 1) : PAIR_NONE; }
    else { d.x = idx - (in_base_ptr ? __ldg(in_base_ptr) : 0u); d.y = has2 ? 0u : PAIR_NONE; }
    desc[o] = d;
}

struct PairDesc {            // one lane's output of one step
    uint32_t p, q;           // as written by k_pair_desc — possibly still in flight: only touch them when the step is issued / computed
    bool valid;              // the output exists (known from the indices alone, never from loaded data)
    FF_DEV bool has2() const { return valid && q != PAIR_NONE; }
};
// descriptor of step j for this lane (dlane = desc + W0 + lane); steps outside [0, nv) do not exist
FF_DEV PairDesc pair_load_desc(int64_t j, uint32_t nv, const uint2* __restrict__ dlane) {
    PairDesc d; d.p = 0; d.q = PAIR_NONE; d.valid = false;
    if (j < 0 || j >= (int64_t)nv) return d;
    const uint2 v = __ldg(dlane + 32 * j);
    d.p = v.x; d.q = v.y; d.valid = true;
    return d;
}
template <bool GATHER>
FF_DEV const uint32_t* pair_src(const PairDesc& d, int which, const uint32_t* __restrict__ records) {
    if (GATHER) return records + (size_t)((which ? d.q : d.p) & 0x7fffffffu) * BASE_WORDS;
    return records + (size_t)(d.p + (uint32_t)which) * DENSE_WORDS;
}
// copies of step operands into ring stage `st` of `chunks` 16-byte chunks per lane: backward (FULL) x1 y1 x2 y2 in 12 chunks,
// forward x1 x2 in 6 chunks
template <bool GATHER, bool FULL>
FF_DEV void pair_issue(const PairDesc& d, uint4* ring, int st, int lane, const uint32_t* __restrict__ records, int chunks) {
    if (d.valid) {
        const uint32_t dst = smem_addr_u32(ring + (size_t)st * (size_t)(chunks * 32) + lane);
        const uint32_t* p = pair_src<GATHER>(d, 0, records);
#pragma unroll
        for (int k = 0; k < (FULL ? 6 : 3); k++) cp_async16(dst + (uint32_t)k * 512u, p + 4 * k);
        if (d.q != PAIR_NONE) {
            const uint32_t* q = pair_src<GATHER>(d, 1, records);
#pragma unroll
            for (int k = 0; k < (FULL ? 6 : 3); k++) cp_async16(dst + (uint32_t)((FULL ? 6 : 3) + k) * 512u, q + 4 * k);
        }
    }
    cp_async_commit();
}
// full classification of one pair from global memory (rare path of the forward pass: equal x, or an x that is 0)
template <bool GATHER>
FF_DEV int pair_classify_global(const PairDesc& d, const uint32_t* __restrict__ records, Fq& den) {
    DensePoint P = load_dense(pair_src<GATHER>(d, 0, records));
    DensePoint Q = load_dense(pair_src<GATHER>(d, 1, records));
    if (GATHER) { if ((d.p >> 31) && !P.inf) P.y = P.y.neg(); if ((d.q >> 31) && !Q.inf) Q.y = Q.y.neg(); }
    return classify_pair(P, Q, true, den);
}

// The shared inversion is a bubble — one warp works while the CTA's other warps wait at the barrier, and CTAs that start
// together reach it together — so (i) the inverting warp computes ONE inverse limb-per-lane (coop_inverse, ff.cuh) instead
// of 32 redundant copies, and (ii) every CTA takes a slot number from a per-SM counter: slot & 3 names the inverting warp, so
// the CTAs resident on an SM invert on different sub-partitions whatever the block → SM mapping is.
// MINB = resident CTAs per SM the kernel is compiled for: 4 (128 registers) or 3 (168 registers, no spills).
template <bool GATHER, int MINB>
__global__ void __launch_bounds__(PAIR_THREADS, MINB) k_pair_level2(const uint32_t* __restrict__ records /* level 0: dense bases / table; above: dense_in */,
                                                         const uint2* __restrict__ desc, const uint32_t* __restrict__ total_ptr,
                                                         uint32_t T, uint32_t* __restrict__ prefix, uint32_t* __restrict__ dense_out,
                                                         uint32_t* __restrict__ sm_slots) {
    extern __shared__ uint4 pair2_smem[];
    __shared__ uint32_t sh_slot;
    uint32_t* sh_inv = reinterpret_cast<uint32_t*>(pair2_smem);                       // 128 × 48 B
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* ring = pair2_smem + PAIR_THREADS * 3 + (size_t)warp * 2 * RING_STAGE_U4;    // this warp's two stages
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        sh_slot = atomicAdd(sm_slots + (smid & 255u), 1u);
    }
    __syncthreads();
    const int inv_warp = (int)(sh_slot & 3u);
    const uint32_t total = __ldg(total_ptr);
    const uint64_t w0_64 = ((uint64_t)blockIdx.x * (PAIR_THREADS / 32) + (uint32_t)warp) * 32ull * T + (uint32_t)lane;   // this lane's first output
    uint32_t nv = 0;                                                                   // steps that exist for this lane
    if (w0_64 < total) { const uint64_t left = (total - w0_64 + 31) / 32; nv = left < T ? (uint32_t)left : T; }
    const uint2* dlane = desc + w0_64;
    uint32_t* plane = prefix + w0_64 * 12;                                             // prefix of step j at plane + j·32·12
    uint32_t* olane = dense_out + w0_64 * DENSE_WORDS;

    // ---------------- forward: running product of the denominators ----------------
    // A forward step is ONE multiplication, far shorter than a gathered DRAM access: the x-only operands (96 B per lane) fit
    // FOUR ring stages where the backward pass has two, so the copies run three steps ahead of the multiplier
    // (ncu, round 2: with one step of lead the forward loop held 21 % of the kernel's stall samples for 4 % of its instructions).
    Fq run = Fq::one();
    {
        constexpr int PF = 3;                                                   // steps of lead; PF + 1 stages of 6 chunks
        PairDesc q0 = pair_load_desc(0, nv, dlane), q1 = pair_load_desc(1, nv, dlane), q2 = pair_load_desc(2, nv, dlane);
        pair_issue<GATHER, false>(q0, ring, 0, lane, records, 6);
        pair_issue<GATHER, false>(q1, ring, 1, lane, records, 6);
        pair_issue<GATHER, false>(q2, ring, 2, lane, records, 6);
        PairDesc ahead = pair_load_desc(PF, nv, dlane);
        for (uint32_t j = 0; j < T; j++) {
            pair_issue<GATHER, false>(ahead, ring, (int)((j + PF) & 3u), lane, records, 6);       // step j+3 (descriptor loaded a step ago)
            const PairDesc cur = q0;
            q0 = q1; q1 = q2; q2 = ahead;
            ahead = pair_load_desc((int64_t)j + PF + 1, nv, dlane);                              // not touched until the next iteration
            cp_async_wait_3();
            Fq d = Fq::one();
            if (cur.has2()) {
                const uint4* slot_p = ring + (size_t)(j & 3u) * (6 * 32) + lane;
                Fq x1 = ring_fq(slot_p), x2 = ring_fq(slot_p + 3 * 32);
                if (x1 == x2 || x1.is_zero() || x2.is_zero()) {
                    Fq den;
                    if (pair_classify_global<GATHER>(cur, records, den) >= PAIR_ADD) d = den;
                } else {
                    d = x2 - x1;
                }
            }
            run = run * d;
            if (cur.valid) run.store(plane + (size_t)j * (32 * 12));
        }
        cp_async_wait_0();
    }
    Fq inv = cta_shared_inverse_by(run, sh_inv, inv_warp);
    // ---------------- backward: one inverse per pair, then the affine addition ----------------
    {
        PairDesc cur = pair_load_desc((int64_t)T - 1, nv, dlane);
        pair_issue<GATHER, true>(cur, ring, 0, lane, records, 12);
        PairDesc nxt = pair_load_desc((int64_t)T - 2, nv, dlane);
        for (uint32_t k = 0; k < T; k++) {
            const uint32_t j = T - 1 - k;
            pair_issue<GATHER, true>(nxt, ring, (int)((k + 1) & 1u), lane, records, 12);
            PairDesc nn = pair_load_desc((int64_t)j - 2, nv, dlane);
            Fq pf = Fq::one();
            if (cur.valid && j != 0) pf = Fq::load(plane + (size_t)(j - 1) * (32 * 12));     // behind the first multiplication
            cp_async_wait_1();
            const uint4* slot_p = ring + (size_t)(k & 1u) * RING_STAGE_U4 + lane;
            // classify (same decisions as the forward pass)
            int kind = PAIR_COPY1;
            Fq d = Fq::one(), num = Fq::zero();
            const bool has2 = cur.has2();
            const bool negP = GATHER && (cur.p >> 31), negQ = GATHER && has2 && (cur.q >> 31);
            if (has2) {
                Fq x1 = ring_fq(slot_p), x2 = ring_fq(slot_p + 6 * 32);
                if (x1 == x2 || x1.is_zero() || x2.is_zero()) {
                    DensePoint P, Q;
                    P.x = x1; P.y = ring_fq(slot_p + 3 * 32); P.inf = P.x.is_zero() && P.y.is_zero();
                    Q.x = x2; Q.y = ring_fq(slot_p + 9 * 32); Q.inf = Q.x.is_zero() && Q.y.is_zero();
                    if (negP && !P.inf) P.y = P.y.neg();
                    if (negQ && !Q.inf) Q.y = Q.y.neg();
                    Fq den;
                    kind = classify_pair(P, Q, true, den);
                    if (kind >= PAIR_ADD) d = den;
                    if (kind == PAIR_ADD) num = Q.y - P.y;
                    else if (kind == PAIR_DBL) { Fq xx = P.x.sqr(); num = xx.dbl() + xx; }
                } else {
                    kind = PAIR_ADD;
                    d = x2 - x1;
                    Fq y1 = ring_fq(slot_p + 3 * 32), y2 = ring_fq(slot_p + 9 * 32);
                    // (±y2) − (±y1) without negating first: one subtraction or one addition, then at most one negation
                    if (negP == negQ) { num = y2 - y1; if (negP) num = num.neg(); }
                    else { num = y2 + y1; if (negQ) num = num.neg(); }
                }
            }
            Fq inv_next = inv * d;
            Fq inv_d = (j != 0) ? inv * pf : inv;
            inv = inv_next;
            Fq lambda = num * inv_d;
            Fq x3 = lambda.sqr();
            {
                Fq x1 = ring_fq(slot_p), x2 = has2 ? ring_fq(slot_p + 6 * 32) : x1;
                x3 = x3 - x1 - x2;
                Fq t = x1 - x3;
                Fq y3 = lambda * t;
                if (cur.valid) {
                    DensePoint R;
                    if (kind >= PAIR_ADD) {
                        Fq y1 = ring_fq(slot_p + 3 * 32);
                        R.x = x3; R.y = negP ? y3 + y1 : y3 - y1; R.inf = false;           // y3 − (±y1)
                    } else if (kind == PAIR_INF) {
                        R.inf = true; R.x = Fq::zero(); R.y = Fq::zero();
                    } else {
                        const int c0 = (kind == PAIR_COPY2) ? 6 : 0;
                        R.x = ring_fq(slot_p + c0 * 32); R.y = ring_fq(slot_p + (c0 + 3) * 32);
                        R.inf = R.x.is_zero() && R.y.is_zero();
                        if (((kind == PAIR_COPY2) ? negQ : negP) && !R.inf) R.y = R.y.neg();
                    }
                    store_dense(olane + (size_t)j * (32 * DENSE_WORDS), R);
                }
            }
            cur = nxt; nxt = nn;
        }
        cp_async_wait_0();
    }
}

__global__ void k_halve_counts(const uint32_t* __restrict__ off_in, uint32_t* __restrict__ cnt_out, uint32_t total_buckets) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_buckets) cnt_out[i] = (off_in[i + 1] - off_in[i] + 1u) >> 1;
    else if (i == total_buckets) cnt_out[i] = 0;
}
__global__ void k_items_from_offsets(const uint32_t* __restrict__ off, uint32_t* __restrict__ items, uint32_t total_buckets, uint32_t cap,
                                     uint32_t* __restrict__ hot, uint32_t hot_max, uint32_t keep) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_buckets) {
        const uint32_t it = (off[i + 1] - off[i] + cap - 1u) / cap;
        items[i] = it;
        if (hot != nullptr && it > keep) { const uint32_t pos = atomicAdd(hot, 1u); if (pos < hot_max) hot[1u + pos] = i; }
    } else if (i == total_buckets) items[i] = 0;
}

// XYZZ accumulation of what the pair levels left: contiguous dense points, no gather, no signs.
__global__ void __launch_bounds__(MSM_ACC_THREADS, MSM_ACC_MINBLOCKS) k_bucket_accumulate_dense(
    const uint32_t* __restrict__ dense, const uint32_t* __restrict__ bucket_start, const uint32_t* __restrict__ item_start,
    uint32_t total_buckets, uint32_t cap, uint32_t* __restrict__ partial) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t total_items = item_start[total_buckets];
    if (t >= total_items) return;
    uint32_t lo = 0, hi = total_buckets;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (item_start[mid] <= t) lo = mid; else hi = mid; }
    uint32_t wb = lo, seg = t - item_start[wb];
    uint32_t b0 = bucket_start[wb], b1 = bucket_start[wb + 1];
    uint32_t s0 = b0 + seg * cap, s1 = s0 + cap < b1 ? s0 + cap : b1;
    XYZZ acc = XYZZ::infinity();
    DensePoint p = load_dense(dense + (size_t)s0 * DENSE_WORDS);
    for (uint32_t k = s0; k < s1; k++) {
        AffinePoint cur; cur.x = p.x; cur.y = p.y; cur.inf = p.inf;
        if (k + 1 < s1) p = load_dense(dense + (size_t)(k + 1) * DENSE_WORDS);
        acc.add_affine(cur, false);
    }
    acc.store(partial + (size_t)t * XYZZ_WORDS);
}

// Hot buckets (all scalars equal; or the top signed-digit window, which holds only the carry and therefore
// puts ~n/2 points into ONE bucket) produce thousands of item partials for one bucket.  They are folded
// 32 at a time by as many threads as there are groups, pass after pass, until every bucket has one partial.
__global__ void k_group_counts(const uint32_t* __restrict__ start_in, uint32_t* __restrict__ cnt_out, uint32_t total_buckets) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_buckets) cnt_out[i] = (start_in[i + 1] - start_in[i] + 31u) >> 5;
    else if (i == total_buckets) cnt_out[i] = 0;
}
__global__ void __launch_bounds__(128) k_partial_group_sum(const uint32_t* __restrict__ partial_in, const uint32_t* __restrict__ start_in,
                                                            const uint32_t* __restrict__ start_out, uint32_t total_buckets,
                                                            uint32_t* __restrict__ partial_out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= start_out[total_buckets]) return;
    uint32_t lo = 0, hi = total_buckets;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (start_out[mid] <= t) lo = mid; else hi = mid; }
    uint32_t g = t - start_out[lo];
    uint32_t i0 = start_in[lo] + g * 32u, i1 = start_in[lo + 1];
    if (i0 + 32u < i1) i1 = i0 + 32u;
    XYZZ s = XYZZ::load(partial_in + (size_t)i0 * XYZZ_WORDS);
    for (uint32_t i = i0 + 1; i < i1; i++) s.add(XYZZ::load(partial_in + (size_t)i * XYZZ_WORDS));
    s.store(partial_out + (size_t)t * XYZZ_WORDS);
}

// rounds of 32:1 folds a bucket with `cnt` item partials takes until at most `keep` are left
FF_DEV uint32_t fold_rounds_of(uint32_t cnt, uint32_t& final_cnt, uint32_t keep = 32u) {
    uint32_t r = 0;
    while (cnt > keep) { cnt = (cnt + 31u) >> 5; r++; }
    final_cnt = cnt;
    return r;
}

// Σ of a bucket's item partials
FF_DEV XYZZ bucket_sum(const uint32_t* __restrict__ partial, const uint32_t* __restrict__ item_start, uint32_t wb) {
    uint32_t i0 = item_start[wb], i1 = item_start[wb + 1];
    XYZZ s = XYZZ::infinity();
    for (uint32_t i = i0; i < i1; i++) {
        if (i == i0) s = XYZZ::load(partial + (size_t)i * XYZZ_WORDS);
        else s.add(XYZZ::load(partial + (size_t)i * XYZZ_WORDS));
    }
    return s;
}

// Thread j of window w owns bucket values [lo, hi] = [j·K + 1, (j+1)·K]:
//   running = Σ S_b ; acc = Σ (b − lo + 1)·S_b   (top-down running sum, batched.rs:356-361)
//   out = acc + (lo − 1)·running = Σ b·S_b over the chunk.
// PAIRS: the chunk's entry for the quad-lane combine levels instead — out[2t] = Σ (b − lo + 1)·S_b, out[2t + 1] = Σ S_b — which
// apply the chunk's offset as 8:1 weighted folds (k_combine_level_quad).  The (lo − 1)·running product costs this thread ≈ 24
// more dependent point operations (16 doublings + the additions of lo's bits) on top of its 32: 43 % of the kernel at 2^24 points.
// folds > 0 (PAIRS only): hot buckets were folded down to ONE partial by k_fold_hot_quad (keep = 1); that partial sits at the bucket's
// offset in `partial` or, after an odd number of rounds, in `partial_b`.
template <bool PAIRS>
__global__ void __launch_bounds__(128) k_bucket_reduce(const uint32_t* __restrict__ partial, const uint32_t* __restrict__ item_start,
                                                        uint32_t nbuckets, uint32_t chunk, uint32_t chunks_per_window,
                                                        uint32_t nwin, uint32_t* __restrict__ out,
                                                        const uint32_t* __restrict__ partial_b = nullptr, int folds = 0) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks_per_window * nwin) return;
    uint32_t w = t / chunks_per_window, j = t % chunks_per_window;
    uint32_t lo = j * chunk, hi = lo + chunk;                 // 0-based bucket indices [lo, hi)
    XYZZ running = XYZZ::infinity(), acc = XYZZ::infinity();
    for (uint32_t b = hi; b-- > lo;) {
        XYZZ s;
        if (PAIRS && folds > 0) {
            const uint32_t wb = w * nbuckets + b, i0 = item_start[wb];
            uint32_t cnt = item_start[wb + 1] - i0;
            const uint32_t r = fold_rounds_of(cnt, cnt, 1u);
            s = cnt ? XYZZ::load(((r & 1u) ? partial_b : partial) + (size_t)i0 * XYZZ_WORDS) : XYZZ::infinity();
        } else
            s = bucket_sum(partial, item_start, w * nbuckets + b);
        running.add(s);
        acc.add(running);
    }
    if (PAIRS) {
        acc.store(out + (size_t)t * 2 * XYZZ_WORDS);
        running.store(out + ((size_t)t * 2 + 1) * XYZZ_WORDS);
        return;
    }
    if (lo != 0u) acc.add(running.mul_u32(lo));               // bucket value of index lo is lo+1 ⇒ (lo+1−1)·running
    acc.store(out + (size_t)t * XYZZ_WORDS);
}

// out[j] = Σ in[j·group .. min((j+1)·group, per_row)) for each of `rows` independent rows.
__global__ void __launch_bounds__(128) k_group_sum(const uint32_t* __restrict__ in, uint32_t per_row, uint32_t group,
                                                    uint32_t out_per_row, uint32_t rows, uint32_t* __restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= out_per_row * rows) return;
    uint32_t r = t / out_per_row, j = t % out_per_row;
    uint32_t i0 = j * group, i1 = i0 + group < per_row ? i0 + group : per_row;
    XYZZ s = XYZZ::infinity();
    for (uint32_t i = i0; i < i1; i++) s.add(XYZZ::load(in + ((size_t)r * per_row + i) * XYZZ_WORDS));
    s.store(out + (size_t)t * XYZZ_WORDS);
}

// =================================================================================================
// Latency path for small MSMs (≤ 2^18 points: a few thousand buckets, every kernel a handful of CTAs).  A lone thread needs
// ≈ 0.55 µs per Fq multiplication, so a chain of 16 mixed additions per work item, 32 + 24 per reduction chunk and 8 per tree
// level added up to > 1 ms of pure dependency latency at 2^12–2^16 points.  Here the chains are cut with warp shuffles:
//   * k_bucket_accumulate_g8 — EIGHT lanes per work item: each lane adds every 8th entry, then a 3-step shuffle butterfly;
//   * k_bucket_reduce_warp   — one warp per 32 buckets: a 5-step suffix scan gives the running sums Σ_{b' ≥ b} S_b', a 5-step
//                              reduction of those gives Σ (b − lo + 1)·S_b (the reference's running-sum trick,
//                              batched.rs:356-361, as a parallel scan);
//   * k_window_combine_warp  — one warp per bucket set folds its ≤ 32 chunk results: Σ acc_j + 32·Σ j·run_j, the weighted sum
//                              again as scan + reduction, the factor 32 as five doublings.
// =================================================================================================
FF_DEV XYZZ shfl_xor_xyzz(const XYZZ& a, int m) {
    XYZZ r;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        r.X.v[j] = __shfl_xor_sync(0xffffffffu, a.X.v[j], m); r.Y.v[j] = __shfl_xor_sync(0xffffffffu, a.Y.v[j], m);
        r.ZZ.v[j] = __shfl_xor_sync(0xffffffffu, a.ZZ.v[j], m); r.ZZZ.v[j] = __shfl_xor_sync(0xffffffffu, a.ZZZ.v[j], m);
    }
    return r;
}
FF_DEV XYZZ shfl_down_xyzz(const XYZZ& a, int d) {
    XYZZ r;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        r.X.v[j] = __shfl_down_sync(0xffffffffu, a.X.v[j], d); r.Y.v[j] = __shfl_down_sync(0xffffffffu, a.Y.v[j], d);
        r.ZZ.v[j] = __shfl_down_sync(0xffffffffu, a.ZZ.v[j], d); r.ZZZ.v[j] = __shfl_down_sync(0xffffffffu, a.ZZZ.v[j], d);
    }
    return r;
}
// every lane ends with Σ over the warp (butterfly: the two partners of a step compute the same sum)
FF_DEV XYZZ warp_sum_xyzz(XYZZ a) {
#pragma unroll 1
    for (int m = 16; m >= 1; m >>= 1) { XYZZ o = shfl_xor_xyzz(a, m); a.add(o); }
    return a;
}
// lane l ends with Σ_{l' ≥ l} a_l'
FF_DEV XYZZ warp_suffix_scan_xyzz(XYZZ a, int lane) {
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) { XYZZ o = shfl_down_xyzz(a, d); if (lane + d < 32) a.add(o); }
    return a;
}

static constexpr int ACC_G = 8;
__global__ void __launch_bounds__(128, 4) k_bucket_accumulate_g8(const uint32_t* __restrict__ records, const uint32_t* __restrict__ sorted,
                                                                 const uint32_t* __restrict__ bucket_start, const uint32_t* __restrict__ item_start,
                                                                 uint32_t total_buckets, uint32_t cap, uint32_t* __restrict__ partial) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, item = t / ACC_G, sub = t % ACC_G;
    const bool valid = item < item_start[total_buckets];
    XYZZ acc = XYZZ::infinity();
    if (valid) {
        uint32_t lo = 0, hi = total_buckets;          // item_start[lo] <= item < item_start[hi]
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (item_start[mid] <= item) lo = mid; else hi = mid; }
        const uint32_t seg = item - item_start[lo];
        const uint32_t b0 = bucket_start[lo], b1 = bucket_start[lo + 1];
        const uint32_t s0 = b0 + seg * cap, s1 = s0 + cap < b1 ? s0 + cap : b1;
        uint32_t k = s0 + sub;
        if (k < s1) {
            uint32_t e = sorted[k];
            AffinePoint p = load_record(records, e & 0x7fffffffu);
            for (; k < s1; k += ACC_G) {
                const uint32_t e_cur = e;
                const AffinePoint p_cur = p;
                if (k + ACC_G < s1) { e = sorted[k + ACC_G]; p = load_record(records, e & 0x7fffffffu); }
                acc.add_affine(p_cur, (e_cur >> 31) != 0u);
            }
        }
    }
#pragma unroll 1
    for (int m = ACC_G / 2; m >= 1; m >>= 1) { XYZZ o = shfl_xor_xyzz(acc, m); acc.add(o); }
    if (valid && sub == 0) acc.store(partial + (size_t)item * XYZZ_WORDS);
}

// out[(set·chunks + chunk)·2] = Σ_l (l + 1)·S_{32·chunk + l},  out[… + 1] = Σ_l S_{32·chunk + l}
__global__ void __launch_bounds__(128) k_bucket_reduce_warp(const uint32_t* __restrict__ partial, const uint32_t* __restrict__ item_start,
                                                            uint32_t nbuckets, uint32_t chunks, uint32_t nsets, uint32_t* __restrict__ out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= nsets * chunks) return;                                 // whole warps only
    const uint32_t set = warp / chunks, ch = warp % chunks, b = ch * 32u + lane;
    XYZZ s = XYZZ::infinity();
    if (b < nbuckets) s = bucket_sum(partial, item_start, set * nbuckets + b);
    const XYZZ run = warp_suffix_scan_xyzz(s, (int)lane);
    const XYZZ acc = warp_sum_xyzz(run);
    if (lane == 0) {
        acc.store(out + (size_t)warp * 2 * XYZZ_WORDS);
        run.store(out + ((size_t)warp * 2 + 1) * XYZZ_WORDS);
    }
}
// window sum = Σ_j acc_j + 32·Σ_j j·run_j over the set's chunks (chunks ≤ 32)
__global__ void __launch_bounds__(32) k_window_combine_warp(const uint32_t* __restrict__ in, uint32_t chunks, uint32_t* __restrict__ out) {
    const uint32_t set = blockIdx.x, lane = threadIdx.x;
    XYZZ a = XYZZ::infinity(), r = XYZZ::infinity();
    if (lane < chunks) {
        a = XYZZ::load(in + ((size_t)set * chunks + lane) * 2 * XYZZ_WORDS);
        r = XYZZ::load(in + (((size_t)set * chunks + lane) * 2 + 1) * XYZZ_WORDS);
    }
    XYZZ total = warp_sum_xyzz(a);
    if (chunks > 1) {
        XYZZ suf = warp_suffix_scan_xyzz(r, (int)lane);                 // Σ_{i ≥ j} run_i
        if (lane == 0) suf = XYZZ::infinity();                          // Σ_{j ≥ 1} suffix_j = Σ_j j·run_j
        XYZZ w = warp_sum_xyzz(suf);
        for (int k = 0; k < 5; k++) w.dbl();
        total.add(w);
    }
    if (lane == 0) total.store(out + (size_t)set * XYZZ_WORDS);
}

// =================================================================================================
// The same tail with FOUR LANES PER POINT (quad.cuh): an XYZZ addition costs 4 dependent multiplications instead of 14.
//   * k_bucket_accumulate_q8 — one WARP per work item: quad s adds every 8th entry (mixed additions), 3-step butterfly
//                              over the eight quads (sizes where the whole problem is a few thousand items);
//   * k_bucket_reduce_quad   — one warp per 8 buckets: 3-step suffix scan + 3-step sum give (Σ (l+1)·S_l, Σ S_l);
//   * k_window_combine_quad  — one CTA per bucket set folds the (acc, run) entries 8 at a time, level after level in
//                              shared memory: acc' = Σ acc_s + w·Σ s·run_s, run' = Σ run_s for entries spanning w buckets.
// =================================================================================================
FF_DEV XYZZ load_xyzz_plain(const uint32_t* p) {              // global or shared memory, 16-byte aligned
    XYZZ r; const uint4* q = reinterpret_cast<const uint4*>(p);
    uint32_t w[XYZZ_WORDS];
#pragma unroll
    for (int i = 0; i < XYZZ_WORDS / 4; i++) { const uint4 t = q[i]; w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w; }
#pragma unroll
    for (int j = 0; j < 12; j++) { r.X.v[j] = w[j]; r.Y.v[j] = w[12 + j]; r.ZZ.v[j] = w[24 + j]; r.ZZZ.v[j] = w[36 + j]; }
    return r;
}
FF_DEV void store_xyzz_plain(uint32_t* p, const XYZZ& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    uint32_t w[XYZZ_WORDS];
#pragma unroll
    for (int j = 0; j < 12; j++) { w[j] = a.X.v[j]; w[12 + j] = a.Y.v[j]; w[24 + j] = a.ZZ.v[j]; w[36 + j] = a.ZZZ.v[j]; }
#pragma unroll
    for (int i = 0; i < XYZZ_WORDS / 4; i++) q[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

__global__ void __launch_bounds__(128, 2) k_bucket_accumulate_q8(const uint32_t* __restrict__ records, const uint32_t* __restrict__ sorted,
                                                                 const uint32_t* __restrict__ bucket_start, const uint32_t* __restrict__ item_start,
                                                                 uint32_t total_buckets, uint32_t cap, uint32_t* __restrict__ partial) {
    const uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= item_start[total_buckets]) return;               // whole warps
    const Quad Q = Quad::here();
    uint32_t lo = 0, hi = total_buckets;                         // item_start[lo] <= item < item_start[hi]
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (item_start[mid] <= item) lo = mid; else hi = mid; }
    const uint32_t seg = item - item_start[lo];
    const uint32_t b0 = bucket_start[lo], b1 = bucket_start[lo + 1];
    const uint32_t s0 = b0 + seg * cap, s1 = s0 + cap < b1 ? s0 + cap : b1;
    XYZZ acc = XYZZ::infinity();
    {
        // lockstep over the item (see k_bucket_reduce_quad): every quad makes every trip, ∞ beyond the item's end
        AffinePoint p; p.x = Fq::zero(); p.y = Fq::zero(); p.inf = true;
        uint32_t e = 0;
        uint32_t k = s0 + (uint32_t)Q.slot;
        if (k < s1) { e = sorted[k]; p = load_record(records, e & 0x7fffffffu); }
#pragma unroll 1
        for (uint32_t base = s0; base < s1; base += 8, k += 8) {
            __syncwarp();
            const uint32_t e_cur = e;
            const AffinePoint p_cur = p;
            p.inf = true;
            if (k + 8 < s1) { e = sorted[k + 8]; p = load_record(records, e & 0x7fffffffu); }
            acc = quad_add_affine(acc, p_cur, (e_cur >> 31) != 0u, Q);
        }
    }
    acc = slot_sum(acc, Q);
    if ((threadIdx.x & 31u) == 0u) acc.store(partial + (size_t)item * XYZZ_WORDS);
}

// Hot buckets without scans.  Some buckets are hot by construction: the top window of a 253-bit scalar has a handful of digit
// values (c = 8: 16 buckets share all n points) or is the carry alone (c = 11: one bucket with ≈ 0.14·n points), equal scalars
// add more.  A bucket's item partials stay at partial[item_start[b] …]; a round replaces every 32 consecutive ones by their sum at
// the SAME base offset of the other buffer (count → ⌈count / 32⌉) while the count exceeds 32, so no offsets are recomputed and
// each bucket knows from its own count how many rounds it took part in and which buffer holds its partials.  k_items_per_bucket
// lists the buckets with more than 32 items; warp w works on hot bucket w / 32 and takes every 32nd group of it, quad s of the
// warp adds entries s, s + 8, … of the group.  Rounds launched for the worst case find nothing to do and return.
__global__ void __launch_bounds__(128) k_fold_hot_quad(const uint32_t* __restrict__ in, const uint32_t* __restrict__ item_start,
                                                       const uint32_t* __restrict__ hot, uint32_t hot_max, uint32_t round, uint32_t keep,
                                                       uint32_t* __restrict__ out) {
    uint32_t nhot = hot[0];
    if (nhot > hot_max) nhot = hot_max;
    const Quad Q = Quad::here();
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5, w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint32_t h = w0 >> 5; h < nhot; h += nwarps >> 5) {             // whole warps; nwarps is a multiple of 32
        const uint32_t b = hot[1u + h];
        const uint32_t i0 = item_start[b];
        uint32_t cnt = item_start[b + 1] - i0;
        bool live = true;
        for (uint32_t r = 0; r < round; r++) { if (cnt <= keep) live = false; cnt = (cnt + 31u) >> 5; }
        if (!live || cnt <= keep) continue;                             // this bucket finished in an earlier round
        const uint32_t groups = (cnt + 31u) >> 5;
        for (uint32_t g = w0 & 31u; g < groups; g += 32u) {
            const uint32_t k = g << 5, end = k + 32u < cnt ? k + 32u : cnt;
            XYZZ s = XYZZ::infinity();
#pragma unroll 1
            for (uint32_t i = k; i < end; i += 8u) {                    // lockstep: every quad makes every trip
                __syncwarp();
                XYZZ v = XYZZ::infinity();
                if (i + (uint32_t)Q.slot < end) v = XYZZ::load(in + (size_t)(i0 + i + (uint32_t)Q.slot) * XYZZ_WORDS);
                s = quad_add(s, v, Q);
            }
            s = slot_sum(s, Q);
            if ((threadIdx.x & 31u) == 0u) s.store(out + (size_t)(i0 + g) * XYZZ_WORDS);
        }
    }
}

// out[(set·chunks + chunk)·2] = Σ_l (l + 1)·S_{8·chunk + l},  out[… + 1] = Σ_l S_{8·chunk + l}
// The item partials of bucket b start at item_start[b] in partial_a, or, after an odd number of folds (fold_rounds_of its
// item count), in partial_b.  `folds` = 0: no fold kernel ran (every bucket has one partial or the caller folded already).
__global__ void __launch_bounds__(128) k_bucket_reduce_quad(const uint32_t* __restrict__ partial_a, const uint32_t* __restrict__ partial_b,
                                                            const uint32_t* __restrict__ item_start, int folds,
                                                            uint32_t nbuckets, uint32_t chunks, uint32_t nsets, uint32_t* __restrict__ out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nsets * chunks) return;                                 // whole warps only
    const Quad Q = Quad::here();
    const uint32_t set = warp / chunks, ch = warp % chunks, b = ch * 8u + (uint32_t)Q.slot;
    XYZZ s = XYZZ::infinity();
    {
        // The eight quads walk their buckets in LOCKSTEP (trip count = the largest item count of the warp, ∞ operands beyond a
        // quad's own count, a __syncwarp per trip): quads that leave a loop at different trips are scheduled one after the other
        // from then on, which made each addition cost 8× (36 µs instead of 4.5, profiles/r2u_cap.log).
        uint32_t i0 = 0, cnt = 0;
        const uint32_t* partial = partial_a;
        if (b < nbuckets) {
            const uint32_t wb = set * nbuckets + b;
            i0 = item_start[wb];
            cnt = item_start[wb + 1] - i0;
            if (folds) { if (fold_rounds_of(cnt, cnt) & 1u) partial = partial_b; }
        }
        const uint32_t trips = __reduce_max_sync(0xffffffffu, cnt);
#pragma unroll 1
        for (uint32_t i = 0; i < trips; i++) {
            __syncwarp();
            XYZZ t = XYZZ::infinity();
            if (i < cnt) t = XYZZ::load(partial + (size_t)(i0 + i) * XYZZ_WORDS);
            s = quad_add(s, t, Q);
        }
    }
    const XYZZ run = slot_suffix_scan(s, Q);
    const XYZZ acc = slot_sum(run, Q);
    if ((threadIdx.x & 31u) == 0u) {
        acc.store(out + (size_t)warp * 2 * XYZZ_WORDS);
        run.store(out + ((size_t)warp * 2 + 1) * XYZZ_WORDS);
    }
}

// One 8:1 fold of (acc, run) entries that span 2^lgw buckets each: acc' = Σ_s acc_s + 2^lgw·Σ_s s·run_s, run' = Σ_s run_s.
// Called by one warp with slot s holding entry 8·group + s (∞ beyond the end); every lane returns acc', `run_out` = run'.
FF_DEV XYZZ combine_fold_quad(const XYZZ& a, const XYZZ& r, int lgw, const Quad& Q, XYZZ& run_out) {
    XYZZ A = slot_sum(a, Q);
    const XYZZ suf = slot_suffix_scan(r, Q);                            // Σ_{s' ≥ s} run_s'
    XYZZ W = slot_sum(Q.slot == 0 ? XYZZ::infinity() : suf, Q);         // Σ_{s ≥ 1} suffix_s = Σ_s s·run_s
#pragma unroll 1
    for (int k = 0; k < lgw; k++) W = quad_dbl(W, Q);
    run_out = suf;                                                      // slot 0 holds the total
    return quad_add(A, W, Q);
}
// one level for sets with more than 64 entries: m entries per set → ceil(m / 8), one warp per output entry
__global__ void __launch_bounds__(128) k_combine_level_quad(const uint32_t* __restrict__ in, uint32_t m, uint32_t nsets, int lgw, uint32_t* __restrict__ out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, groups = (m + 7u) / 8u;
    if (warp >= nsets * groups) return;
    const Quad Q = Quad::here();
    const uint32_t set = warp / groups, g = warp % groups, j = g * 8u + (uint32_t)Q.slot;
    XYZZ a = XYZZ::infinity(), r = XYZZ::infinity();
    if (j < m) { a = load_xyzz_plain(in + ((size_t)set * m + j) * 2 * XYZZ_WORDS); r = load_xyzz_plain(in + (((size_t)set * m + j) * 2 + 1) * XYZZ_WORDS); }
    XYZZ run;
    const XYZZ A = combine_fold_quad(a, r, lgw, Q, run);
    if ((threadIdx.x & 31u) == 0u) {
        store_xyzz_plain(out + (size_t)warp * 2 * XYZZ_WORDS, A);
        store_xyzz_plain(out + ((size_t)warp * 2 + 1) * XYZZ_WORDS, run);
    }
}
// The same fold with ONE QUAD per group, its 8 entries walked from the top with running sums (Σ s·run_s = Σ_{s ≥ 1} running after
// entry s): 23 additions per group instead of 9 scan / sum steps that all eight quads of a warp execute — 4× fewer warp
// instructions.  For levels with thousands of groups (4096 → 512 entries per window at 2^24 points: 0.69 → 0.2 ms), where the
// multiplier's throughput matters more than the depth of one group's chain.
__global__ void __launch_bounds__(128) k_combine_level_quadseq(const uint32_t* __restrict__ in, uint32_t m, uint32_t nsets, int lgw, uint32_t* __restrict__ out) {
    const uint32_t quad = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, groups = (m + 7u) / 8u;
    if (quad >= nsets * groups) return;                                 // whole quads; no warp-wide exchange below
    const Quad Q = Quad::here();
    const uint32_t set = quad / groups, g = quad % groups, first = g * 8u;
    const uint32_t cnt = m - first < 8u ? m - first : 8u;
    XYZZ A = XYZZ::infinity(), running = XYZZ::infinity(), W = XYZZ::infinity();
#pragma unroll 1
    for (uint32_t s = cnt; s-- > 0u;) {
        const size_t e = ((size_t)set * m + first + s) * 2;
        A = quad_add(A, load_xyzz_plain(in + e * XYZZ_WORDS), Q);
        running = quad_add(running, load_xyzz_plain(in + (e + 1) * XYZZ_WORDS), Q);
        if (s >= 1u) W = quad_add(W, running, Q);
    }
#pragma unroll 1
    for (int k = 0; k < lgw; k++) W = quad_dbl(W, Q);
    A = quad_add(A, W, Q);
    if (Q.q == 0) {
        store_xyzz_plain(out + (size_t)quad * 2 * XYZZ_WORDS, A);
        store_xyzz_plain(out + ((size_t)quad * 2 + 1) * XYZZ_WORDS, running);
    }
}
// window sum of a set from its m0 ≤ 64 entries (each spanning 2^lgw0 buckets): one CTA per set, 8:1 per level through
// shared memory
static constexpr int COMBINE_QUAD_MAX_WARPS = 8;
__global__ void __launch_bounds__(32 * COMBINE_QUAD_MAX_WARPS) k_window_combine_quad(const uint32_t* __restrict__ in, uint32_t m0, int lgw0, uint32_t* __restrict__ out) {
    __shared__ __align__(16) uint32_t sm[COMBINE_QUAD_MAX_WARPS][2][XYZZ_WORDS];
    const uint32_t set = blockIdx.x, warp = threadIdx.x >> 5;
    const Quad Q = Quad::here();
    const uint32_t* src = in + (size_t)set * m0 * 2 * XYZZ_WORDS;
    uint32_t m = m0;
    int lgw = lgw0;
    for (;;) {                                                          // m ≤ 64 → ≤ 8 → 1
        const uint32_t groups = (m + 7u) / 8u;
        XYZZ A, run;
        if (warp < groups) {
            const uint32_t j = warp * 8u + (uint32_t)Q.slot;
            XYZZ a = XYZZ::infinity(), r = XYZZ::infinity();
            if (j < m) { a = load_xyzz_plain(src + (size_t)j * 2 * XYZZ_WORDS); r = load_xyzz_plain(src + ((size_t)j * 2 + 1) * XYZZ_WORDS); }
            A = combine_fold_quad(a, r, lgw, Q, run);
            if (groups == 1u && (threadIdx.x & 31u) == 0u) A.store(out + (size_t)set * XYZZ_WORDS);
        }
        if (groups == 1u) break;
        __syncthreads();                                                // the previous level's entries have been read
        if (warp < groups && (threadIdx.x & 31u) == 0u) { store_xyzz_plain(&sm[warp][0][0], A); store_xyzz_plain(&sm[warp][1][0], run); }
        __syncthreads();
        src = &sm[0][0][0];
        m = groups; lgw += 3;
    }
}

__global__ void k_xyzz_sum_ranks(const uint32_t* __restrict__ in, int nranks, int count, uint32_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    XYZZ s = XYZZ::load(in + (size_t)i * XYZZ_WORDS);
    for (int r = 1; r < nranks; r++) s.add(XYZZ::load(in + ((size_t)r * count + i) * XYZZ_WORDS));
    s.store(out + (size_t)i * XYZZ_WORDS);
}

int xyzz_sum_ranks_device(uint32_t* d_out, const uint32_t* d_in, int nranks, int count, cudaStream_t stream) {
    k_xyzz_sum_ranks<<<(count + 31) / 32, 32, 0, stream>>>(d_in, nranks, count, d_out);
    count_launch();
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Scratch: one private stream-ordered pool per device (the default pool is left alone) and a byte budget that
// bounds how much MSM scratch is in flight per device.  The FFI is entered concurrently from many rayon workers
// (sonic_pc/mod.rs:186-245): without the gate, a burst of large commitments would each take tens of GB and the
// losers would fail with cudaErrorMemoryAllocation (⇒ silent CPU fallback on the Rust side); with it they queue.
// ---------------------------------------------------------------------------
struct DeviceScratch {
    std::mutex mu;
    std::condition_variable cv;
    cudaMemPool_t pool = nullptr;
    size_t limit = 0, in_use = 0, peak = 0;
    bool ready = false;
};
static DeviceScratch g_scratch[64];

static int scratch_init(DeviceScratch& ds, int dev) {
    if (ds.ready) return 0;
    size_t free_b = 0, total_b = 0;
    cudaError_t e = cudaMemGetInfo(&free_b, &total_b);
    if (e != cudaSuccess) return (int)e;
    size_t limit = total_b / 2;                                       // default: half the device
    if (const char* v = getenv("SNARKVM_B200_SCRATCH_LIMIT_GB")) { long g = atol(v); if (g >= 1) limit = (size_t)g << 30; }
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    if ((e = cudaMemPoolCreate(&ds.pool, &props)) != cudaSuccess) return (int)e;
    uint64_t thr = limit;                                             // keep up to `limit` cached between calls, not more
    cudaMemPoolSetAttribute(ds.pool, cudaMemPoolAttrReleaseThreshold, &thr);
    ds.limit = limit;
    ds.ready = true;
    return 0;
}
// small, ungated allocations (window sums, NTT scratch, polynomial temporaries) from the same private pool
int pool_alloc_raw(void** p, size_t bytes, cudaStream_t stream) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    DeviceScratch& ds = g_scratch[dev & 63];
    {
        std::lock_guard<std::mutex> lock(ds.mu);
        int rc = scratch_init(ds, dev);
        if (rc) return rc;
    }
    return (int)cudaMallocFromPoolAsync(p, bytes ? bytes : 16, ds.pool, stream);
}
struct ScratchLease { DeviceScratch* ds; size_t bytes; };
static void CUDART_CB scratch_release_cb(void* p) {
    ScratchLease* l = (ScratchLease*)p;
    { std::lock_guard<std::mutex> lock(l->ds->mu); l->ds->in_use -= l->bytes; }
    l->ds->cv.notify_all();
    delete l;
}
// blocks until `bytes` fit in the device's budget (a request larger than the whole budget runs alone), then allocates
static int scratch_acquire(void** out, size_t bytes, cudaStream_t stream, DeviceScratch** ds_out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    DeviceScratch& ds = g_scratch[dev & 63];
    {
        std::unique_lock<std::mutex> lock(ds.mu);
        int rc = scratch_init(ds, dev);
        if (rc) return rc;
        ds.cv.wait(lock, [&] { return ds.in_use == 0 || ds.in_use + bytes <= ds.limit; });
        ds.in_use += bytes;
        if (ds.in_use > ds.peak) ds.peak = ds.in_use;
    }
    e = cudaMallocFromPoolAsync(out, bytes, ds.pool, stream);
    if (e != cudaSuccess) {
        { std::lock_guard<std::mutex> lock(ds.mu); ds.in_use -= bytes; }
        ds.cv.notify_all();
        return (int)e;
    }
    *ds_out = &ds;
    return 0;
}
// frees in stream order and returns the bytes to the budget when the stream gets there
static void scratch_release(void* p, size_t bytes, cudaStream_t stream, DeviceScratch* ds) {
    cudaFreeAsync(p, stream);
    ScratchLease* l = new ScratchLease{ds, bytes};
    if (cudaLaunchHostFunc(stream, scratch_release_cb, l) != cudaSuccess) scratch_release_cb(l);
}
int msm_scratch_stats(size_t* limit, size_t* in_use, size_t* peak) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    DeviceScratch& ds = g_scratch[dev & 63];
    std::lock_guard<std::mutex> lock(ds.mu);
    if (limit) *limit = ds.limit;
    if (in_use) *in_use = ds.in_use;
    if (peak) *peak = ds.peak;
    return 0;
}

int msm_set_scratch_limit(size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (bytes == 0) return (int)cudaErrorInvalidValue;
    DeviceScratch& ds = g_scratch[dev & 63];
    {
        std::lock_guard<std::mutex> lock(ds.mu);
        int rc = scratch_init(ds, dev);
        if (rc) return rc;
        ds.limit = bytes;
        ds.peak = ds.in_use;
        uint64_t thr = bytes;
        cudaMemPoolSetAttribute(ds.pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    ds.cv.notify_all();
    return 0;
}

struct Arena {
    uint8_t* base = nullptr;
    size_t off = 0;
    template <class T> T* take(size_t count) {
        T* p = base ? (T*)(base + off) : nullptr;
        off += (count * sizeof(T) + 255) & ~(size_t)255;
        return p;
    }
};

// The general form: `njobs` independent sums (one per committed polynomial), each fed by one or more scalar segments,
// over ONE set of resident bases.  All jobs go through one digit/sort pass keyed by (job, window, bucket), one set of pair
// levels and one reduction; d_window_sums receives njobs × (flat ? 1 : nwin) XYZZ points, job-major.
// table != nullptr: "flat" mode over precomputed tables (record w·table_n + i = 2^{c·w}·P_i): all windows of a job share
// one bucket set, so the pipeline sees ONE set of n·nwin entries per job and writes a single sum per job.
int msm_core(uint32_t* d_window_sums, uint32_t* d_flags, const MsmPlan& plan, const MsmBases* bases, int nbases,
             const uint32_t* table, size_t table_n, const MsmSegment* segs, int nsegs, int njobs, cudaStream_t stream) {
    int rc = 0;
    const bool flat = table != nullptr;
    const uint32_t sets_per_job = flat ? 1u : (uint32_t)plan.nwin;   // bucket sets to reduce per job
    if (njobs < 1 || nsegs < 1 || (!flat && nbases < 1)) return (int)cudaErrorInvalidValue;
    const uint64_t nsets64 = (uint64_t)njobs * sets_per_job;
    const uint64_t TB64 = nsets64 * plan.nbuckets;
    if (TB64 >= (1ull << 31)) return (int)cudaErrorInvalidValue;
    const uint32_t nsets = (uint32_t)nsets64, TB = (uint32_t)TB64;
    size_t total_scalars = 0, total_bases = 0;
    std::vector<size_t> job_n((size_t)njobs, 0);
    for (int i = 0; i < nbases; i++) {
        if (bases[i].stride < 104 || (bases[i].stride & 7)) return (int)cudaErrorInvalidValue;
        total_bases += bases[i].n;
    }
    for (int i = 0; i < nsegs; i++) {
        if (segs[i].job >= (uint32_t)njobs) return (int)cudaErrorInvalidValue;
        if (flat ? (segs[i].base0 != 0 || segs[i].n > table_n) : ((size_t)segs[i].base0 + segs[i].n > total_bases)) return (int)cudaErrorInvalidValue;
        total_scalars += segs[i].n;
        job_n[segs[i].job] += segs[i].n;
    }
    size_t max_job_n = 0;
    for (size_t v : job_n) if (v > max_job_n) max_job_n = v;
    const size_t max_entries = total_scalars * (size_t)plan.nwin;
    if (total_scalars == 0 || total_bases >= (1ull << 31) || max_entries >= (1ull << 32)) return (int)cudaErrorInvalidValue;
    if (flat && table_n * (size_t)plan.nwin >= (1ull << 31)) return (int)cudaErrorInvalidValue;
    const int levels = plan.levels;
    const size_t set_cap = flat ? max_job_n * (size_t)plan.nwin : max_job_n;   // most entries one bucket set (or one bucket) can hold
    // pair levels over plain bases: the sort scatters 96-byte records (k_scatter_records) and every level is dense;
    // tables (flat) and the XYZZ-only path keep the index sort + gather
    bool records = levels > 0;
    if (const char* e = getenv("SNARKVM_B200_MSM_RECORDS")) records = records && atoi(e) != 0;
    if (const char* e = getenv("SNARKVM_B200_MSM_PAIR_V1")) { if (atoi(e) != 0) records = false; }      // the round-1 kernel gathers
    // a scalar segment must lie inside one base array (the record scatter reads its points through one pointer)
    std::vector<const uint8_t*> seg_points((size_t)nsegs, nullptr);
    std::vector<size_t> seg_stride((size_t)nsegs, 0);
    if (records && !flat) {
        for (int i = 0; i < nsegs; i++) {
            size_t off = 0; bool found = segs[i].n == 0;
            for (int k = 0; k < nbases && !found; k++) {
                if (segs[i].base0 >= off && (size_t)segs[i].base0 + segs[i].n <= off + bases[k].n) {
                    seg_points[i] = (const uint8_t*)bases[k].d_points + ((size_t)segs[i].base0 - off) * bases[k].stride;
                    seg_stride[i] = bases[k].stride;
                    found = true;
                }
                off += bases[k].n;
            }
            if (!found) return (int)cudaErrorInvalidValue;
        }
    }

    // Everything after the bucket sort runs per GROUP of whole bucket sets, so the dense scratch of the pair levels
    // (≈ 200 B per entry with the level-0 records) stays inside a budget: 2^24 points → 15 windows in one group (49 GB), a
    // 2^26-point shard four windows at a time.
    size_t budget = (size_t)64 << 30;
    if (const char* e = getenv("SNARKVM_B200_MSM_SCRATCH_GB")) { long v = atol(e); if (v >= 1) budget = (size_t)v << 30; }
    if (const char* e = getenv("SNARKVM_B200_MSM_SCRATCH_MB")) { long v = atol(e); if (v >= 1) budget = (size_t)v << 20; }      // tests: force many groups
    uint32_t gw = nsets;
    if (levels > 0) {
        // per entry: dense_a 48 + dense_b 24 + prefix 24 + descriptors 4, plus the 96-byte level-0 records or the 4-byte index
        size_t per_set = set_cap * (size_t)(records ? 196 : 104) + 1;
        size_t fit = budget / per_set;
        if (fit < 1) fit = 1;
        if (fit < gw) {
            const uint32_t ngroups = (uint32_t)((nsets + fit - 1) / fit);          // equal groups (15 sets in 3 groups: 5 + 5 + 5, not 7 + 7 + 1)
            gw = (nsets + ngroups - 1) / ngroups;
        }
    }
    const uint32_t TBg = gw * plan.nbuckets;                      // buckets of the largest group
    size_t entries_g = set_cap * (size_t)gw;                      // most entries a group can hold
    if (entries_g > max_entries) entries_g = max_entries;
    if (records && entries_g >= 0x7fffffffull) return (int)cudaErrorInvalidValue;      // record positions carry the sign in bit 31
    // small problems take the shuffle-based latency path (see k_bucket_reduce_warp)
    bool warp_reduce = plan.nbuckets <= 1024u;
    // (measured, profiles/r2i_ab.log: the shuffle reduction wins at every size it applies to — 1.05 → 0.5 ms; eight lanes per
    // item win only while the whole problem is a few CTAs — 2^10: 0.20 → 0.10 ms, 2^12 equal, 2^14 and up lose to the butterfly's
    // extra additions)
    bool acc_g8 = levels == 0 && max_entries <= 100000;
    // round 2 (quad.cuh): four lanes per point operation.  One warp per work item (k_bucket_accumulate_q8) while the whole
    // problem is a few thousand buckets — every lane of the warp executes every multiplication, so it costs 3× the
    // multiplier time of the one-thread kernel and only pays while the GPU is mostly idle; the item holds up to 1/16 of a
    // bucket set, so a bucket has ≤ 17 item partials, k_bucket_reduce_quad adds them itself and the 32:1 folds are skipped.
    int quad_path = 1;
    if (const char* e = getenv("SNARKVM_B200_MSM_QUAD")) quad_path = atoi(e);
    // (measured, profiles/r2s_phases_*.log: the warp-per-item kernel wins at 2^8 points — 0.10 → 0.06 ms — and loses from 2^10,
    // 0.105 → 0.134 ms, where the 1400 items no longer fit one wave of 255-register warps)
    bool acc_q8 = quad_path != 0 && warp_reduce && levels == 0 && (size_t)TB + max_entries / 32 <= 1100;
    if (const char* e = getenv("SNARKVM_B200_MSM_WARP_PATH")) { if (atoi(e) == 0) { warp_reduce = false; acc_g8 = false; acc_q8 = false; } }
    bool quad_tail_large = true;                                  // quad-lane combine levels after the per-chunk reduction of large bucket sets
    // (measured, profiles/r2ae_phases.log: total 2^24 90.75 → 89.59 ms, 2^22 28.02 → 27.29, 2^20 9.82 → 9.41; at 2^19 — 4096 buckets per
    // window, 256 chunks — the shorter old chain wins, 7.28 vs 7.43 ms)
    if (plan.nbuckets < 16384u) quad_tail_large = false;
    if (const char* e = getenv("SNARKVM_B200_MSM_QUAD_TAIL")) quad_tail_large = atoi(e) != 0;
    // scan-free 32:1 folds of the hot buckets (a device-side list of those with more than 32 item partials)
    const bool quad_fold = quad_path != 0 && warp_reduce && levels == 0;
    // large bucket sets: the same list-driven folds, down to ONE partial per bucket (a lone thread adds what is left of a bucket in
    // k_bucket_reduce, so nothing may be left), instead of two scan + copy passes over every bucket
    const bool large_hot = quad_path != 0 && quad_tail_large && !warp_reduce;
    const uint32_t hot_keep = large_hot ? 1u : 32u;
    if (acc_q8) acc_g8 = false;
    uint32_t item_cap = plan.cap;                                  // points per work item of the XYZZ accumulation
    if (acc_g8) {                                                  // eight lanes per item: 8 × (4 … 16) points
        size_t per_lane = max_entries / 300000 + 1;
        if (per_lane < 4) per_lane = 4;
        if (per_lane > 16) per_lane = 16;
        item_cap = (uint32_t)per_lane * 8u;
        if (const char* e = getenv("SNARKVM_B200_MSM_CAP")) { long v = atol(e); if (v >= 1) item_cap = (uint32_t)v; }
    }
    if (acc_q8) {
        size_t cap16 = ((set_cap + 15) / 16 + 7) & ~(size_t)7;
        item_cap = (uint32_t)(cap16 < 32 ? 32 : cap16);
    } else if (quad_fold && !acc_g8 && max_entries <= 900000) {
        // latency-bound sizes (2^12 … 2^15 points): 4 … 8 dependent mixed additions per thread instead of 16; the extra item
        // partials of a bucket cost the quad reduction 4 multiplication steps each (profiles/r2u_cap.log: 2^12 0.78 → 0.66 ms
        // with 4, 2^13 0.95 → 0.85 and 2^15 1.26 → 1.19 with 8)
        item_cap = max_entries <= 150000 ? 4 : 8;
        if (const char* e = getenv("SNARKVM_B200_MSM_CAP")) { long v = atol(e); if (v >= 1) item_cap = (uint32_t)v; }
    }
    const bool small_cap = quad_fold && !acc_g8 && !acc_q8 && item_cap != plan.cap;
    const size_t max_items = (size_t)TBg + entries_g / (item_cap < plan.cap ? item_cap : plan.cap) + 1;
    const size_t hot_max = max_items / 2 + 1;                         // buckets with more than `hot_keep` (1 or 32) item partials
    const size_t dense_cap_a = entries_g / 2 + TBg + 1, dense_cap_b = entries_g / 4 + 2 * (size_t)TBg + 1;

    size_t pair_waves = 0;                       // 0 = fewest whole waves with T ≤ 1024 outputs per lane
    if (const char* e = getenv("SNARKVM_B200_MSM_PAIR_WAVES")) { long v = atol(e); if (v >= 1) pair_waves = (size_t)v; }
    bool pair_v1 = false;                        // A/B switch: the round-1 thread-contiguous pair level
    if (const char* e = getenv("SNARKVM_B200_MSM_PAIR_V1")) pair_v1 = atoi(e) != 0;
    int pair_minb = 4;                           // resident CTAs per SM the pair kernel is built for (4 × 128 regs or 3 × 168 regs)
    if (const char* e = getenv("SNARKVM_B200_MSM_PAIR_MINB")) { int v = atoi(e); if (v == 3 || v == 4) pair_minb = v; }
    int sm_count = 148;
    {
        static std::once_flag smem_once[64];
        int dev = 0; cudaGetDevice(&dev);
        std::call_once(smem_once[dev & 63], [] {
            cudaFuncSetAttribute(k_pair_level2<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR2_SMEM);
            cudaFuncSetAttribute(k_pair_level2<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR2_SMEM);
            cudaFuncSetAttribute(k_pair_level2<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR2_SMEM);
            cudaFuncSetAttribute(k_pair_level2<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR2_SMEM);
        });
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
    }
    // The reduction tail is a chain of dependent point additions per thread (≈ 16 µs each for a lone warp): short chunks
    // and a narrow (8:1) tree keep that chain short — what matters at 2^16–2^20 points, where the tail is 20–50 % of the call.
    const uint32_t chunk = plan.nbuckets < 16u ? plan.nbuckets : 16u;
    const uint32_t tree = 8;
    const uint32_t chunks_per_set = plan.nbuckets / chunk;

    size_t cub_bytes = 0;
    if (cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(TB + 1), stream) != cudaSuccess) return (int)cudaErrorUnknown;
    if (cub_bytes < 16) cub_bytes = 16;

    // ---- one scratch block, carved up ----
    uint32_t *hist, *bucket_start, *cursors, *items, *item_start, *items2, *sorted, *partial, *partial2, *red_a, *red_b;
    uint32_t *off_a = nullptr, *off_b = nullptr, *dense_a = nullptr, *dense_b = nullptr, *prefix = nullptr, *dense_bases = nullptr, *sm_slots = nullptr;
    uint32_t *dense0 = nullptr, *cnt_tmp = nullptr, *hot_dev = nullptr;
    uint2* desc = nullptr;
    uint8_t* cub_tmp;
    Arena ar;
    auto layout = [&](Arena& a) {
        hist = a.take<uint32_t>((size_t)TB + 1);
        bucket_start = a.take<uint32_t>((size_t)TB + 1);
        cursors = a.take<uint32_t>((size_t)TB + 1);
        items = a.take<uint32_t>((size_t)TBg + 1);
        item_start = a.take<uint32_t>((size_t)TBg + 1);
        items2 = a.take<uint32_t>((size_t)TBg + 1);
        sorted = records ? nullptr : a.take<uint32_t>(max_entries);
        cnt_tmp = a.take<uint32_t>((size_t)TBg + 1);
        partial = a.take<uint32_t>(max_items * XYZZ_WORDS);
        partial2 = a.take<uint32_t>(((quad_fold || large_hot) ? max_items : (size_t)TBg + max_items / 32 + 2) * XYZZ_WORDS);      // quad folds keep the item layout
        hot_dev = a.take<uint32_t>(hot_max + 2);
        red_a = a.take<uint32_t>((size_t)gw * (2 * chunks_per_set > 2 * ((plan.nbuckets + 7u) / 8u) ? 2 * chunks_per_set : 2 * ((plan.nbuckets + 7u) / 8u)) * XYZZ_WORDS);
        red_b = a.take<uint32_t>((size_t)gw * (chunks_per_set / tree + 1 > 2 * ((plan.nbuckets + 63u) / 64u) + 2 ? chunks_per_set / tree + 1 : 2 * ((plan.nbuckets + 63u) / 64u) + 2) * XYZZ_WORDS);
        cub_tmp = a.take<uint8_t>(cub_bytes);
        if (!flat && !records) dense_bases = a.take<uint32_t>(total_bases * (size_t)BASE_WORDS);
        if (records) dense0 = a.take<uint32_t>(entries_g * (size_t)DENSE_WORDS);
        if (levels > 0) {
            off_a = a.take<uint32_t>((size_t)TBg + 1);
            off_b = a.take<uint32_t>((size_t)TBg + 1);
            dense_a = a.take<uint32_t>(dense_cap_a * DENSE_WORDS);
            if (levels > 1) dense_b = a.take<uint32_t>(dense_cap_b * DENSE_WORDS);
            prefix = a.take<uint32_t>(dense_cap_a * 12);
            desc = a.take<uint2>(dense_cap_a);
            sm_slots = a.take<uint32_t>(256);
        }
    };
    layout(ar);
    const size_t scratch_bytes = ar.off;
    void* block = nullptr;
    DeviceScratch* ds = nullptr;
    if ((rc = scratch_acquire(&block, scratch_bytes, stream, &ds)) != 0) return rc;
    ar.base = (uint8_t*)block; ar.off = 0;
    layout(ar);

    CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)(TB + 1) * 4, stream));
    if (sm_slots) CUDA_TRY(cudaMemsetAsync(sm_slots, 0, 256 * 4, stream));
    if (quad_fold) CUDA_TRY(cudaMemsetAsync(hot_dev, 0, 4, stream));     // (large_hot: reset per window group below)
    {
        // ---- bucket sort of all jobs and windows: histogram → offsets → scatter ----
        {
            ProfScope sort_scope(PROF_MSM_SORT, stream);
            for (int pass = 0; pass < (records ? 1 : 2); pass++) {
                for (int i = 0; i < nsegs; i++) {
                    const MsmSegment& sg = segs[i];
                    if (sg.n == 0) continue;
                    const unsigned grid = (unsigned)((sg.n + 255) / 256);
                    const uint32_t slot_base = sg.job * sets_per_job * plan.nbuckets;
                    const uint32_t* sc = (const uint32_t*)sg.d_scalars;
                    const size_t fl = flat ? table_n : 0;
                    if (pass == 0) {
                        if (sg.mont) k_digits<false, true><<<grid, 256, 0, stream>>>(sc, sg.n, plan.c, plan.nwin, plan.nbuckets, hist, nullptr, fl, slot_base, sg.base0, d_flags);
                        else k_digits<false, false><<<grid, 256, 0, stream>>>(sc, sg.n, plan.c, plan.nwin, plan.nbuckets, hist, nullptr, fl, slot_base, sg.base0, d_flags);
                    } else {
                        if (sg.mont) k_digits<true, true><<<grid, 256, 0, stream>>>(sc, sg.n, plan.c, plan.nwin, plan.nbuckets, cursors, sorted, fl, slot_base, sg.base0, d_flags);
                        else k_digits<true, false><<<grid, 256, 0, stream>>>(sc, sg.n, plan.c, plan.nwin, plan.nbuckets, cursors, sorted, fl, slot_base, sg.base0, d_flags);
                    }
                    count_launch();
                }
                if (pass == 0) {
                    CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, hist, bucket_start, (int)(TB + 1), stream));
                    CUDA_TRY(cudaMemcpyAsync(cursors, bucket_start, (size_t)(TB + 1) * 4, cudaMemcpyDeviceToDevice, stream));
                    count_launch(2);
                }
            }
        }
        const uint32_t* gather_src = flat ? table : dense_bases;
        if (!flat && !records) {
            ProfScope acc_scope(PROF_MSM_ACCUMULATE, stream);
            size_t at = 0;
            for (int i = 0; i < nbases; i++) {
                if (bases[i].n == 0) continue;
                k_densify_bases<<<(unsigned)((bases[i].n + 255) / 256), 256, 0, stream>>>((const uint8_t*)bases[i].d_points, bases[i].stride, bases[i].n,
                                                                                        dense_bases + at * BASE_WORDS);
                count_launch();
                at += bases[i].n;
            }
        }
        for (uint32_t w0 = 0; w0 < nsets; w0 += gw) {
            const uint32_t wn = nsets - w0 < gw ? nsets - w0 : gw;                                // bucket sets in this group
            const uint32_t tb = wn * plan.nbuckets;
            const uint32_t* bs = bucket_start + (size_t)w0 * plan.nbuckets;                       // tb + 1 absolute offsets into `sorted`
            size_t entries = set_cap * (size_t)wn;                                                // bound on the group's entries
            if (entries > max_entries) entries = max_entries;
            if (large_hot) CUDA_TRY(cudaMemsetAsync(hot_dev, 0, 4, stream));
            size_t items_bound = 1;                                                                // ≥ item count of any single bucket
            size_t items_launched = 1;                                                             // ≥ total item count of the group
            const uint32_t* final_partial = nullptr;
            const uint32_t* final_start = nullptr;
            if (levels == 0) {
                const uint32_t cap = (acc_g8 || acc_q8 || small_cap) ? item_cap : plan.cap;
                k_items_per_bucket<<<(tb + 256) / 256, 256, 0, stream>>>(hist + (size_t)w0 * plan.nbuckets, items, tb, cap, (quad_fold || large_hot) ? hot_dev : nullptr, (uint32_t)hot_max, hot_keep);
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, items, item_start, (int)(tb + 1), stream));
                count_launch(2);
                const size_t group_items = (size_t)tb + entries / cap + 1;
                items_bound = set_cap / cap + 1;
                items_launched = group_items;
                ProfScope acc_scope(PROF_MSM_ACCUMULATE, stream);
                if (acc_q8) {
                    k_bucket_accumulate_q8<<<(unsigned)((group_items * 32 + 127) / 128), 128, 0, stream>>>(gather_src, sorted, bs, item_start, tb, cap, partial);
                    items_bound = 1;                           // ≤ 17 partials per bucket: summed by k_bucket_reduce_quad, no folds
                } else if (acc_g8)
                    k_bucket_accumulate_g8<<<(unsigned)((group_items * ACC_G + 127) / 128), 128, 0, stream>>>(gather_src, sorted, bs, item_start, tb, cap, partial);
                else
                    k_bucket_accumulate<<<(unsigned)((group_items + MSM_ACC_THREADS - 1) / MSM_ACC_THREADS), MSM_ACC_THREADS, 0, stream>>>(
                        gather_src, sorted, bs, item_start, tb, cap, partial);
                count_launch();
            } else {
                if (records) {
                    // this group's records: every segment re-derives its digits and emits the windows that fall into the group
                    ProfScope sort_scope(PROF_MSM_SORT, stream);
                    for (int i = 0; i < nsegs; i++) {
                        const MsmSegment& sg = segs[i];
                        if (sg.n == 0) continue;
                        // windows of this segment inside the group: its job's bucket sets are [job·sets, (job+1)·sets)
                        const int64_t first = (int64_t)sg.job * sets_per_job;
                        int w_lo, w_hi;
                        if (flat) {                                                                   // one set per job: all windows or none
                            if (first < (int64_t)w0 || first >= (int64_t)w0 + wn) continue;
                            w_lo = 0; w_hi = plan.nwin;
                        } else {
                            const int64_t lo_w = (int64_t)w0 - first, hi_w = (int64_t)w0 + wn - first;
                            w_lo = lo_w < 0 ? 0 : (int)lo_w; w_hi = hi_w > plan.nwin ? plan.nwin : (int)hi_w;
                            if (w_lo >= w_hi) continue;
                        }
                        const unsigned grid = (unsigned)((sg.n + 255) / 256);
                        const uint32_t slot_base = sg.job * sets_per_job * plan.nbuckets;
                        const uint32_t* sc = (const uint32_t*)sg.d_scalars;
                        if (flat) {
                            if (sg.mont) k_scatter_records<true, true><<<grid, 256, 0, stream>>>(sc, sg.n, nullptr, 0, table, table_n, plan.c, plan.nbuckets, cursors, slot_base, w_lo, w_hi, bs, (uint4*)dense0);
                            else k_scatter_records<false, true><<<grid, 256, 0, stream>>>(sc, sg.n, nullptr, 0, table, table_n, plan.c, plan.nbuckets, cursors, slot_base, w_lo, w_hi, bs, (uint4*)dense0);
                        } else {
                            if (sg.mont) k_scatter_records<true, false><<<grid, 256, 0, stream>>>(sc, sg.n, seg_points[i], seg_stride[i], nullptr, 0, plan.c, plan.nbuckets, cursors, slot_base, w_lo, w_hi, bs, (uint4*)dense0);
                            else k_scatter_records<false, false><<<grid, 256, 0, stream>>>(sc, sg.n, seg_points[i], seg_stride[i], nullptr, 0, plan.c, plan.nbuckets, cursors, slot_base, w_lo, w_hi, bs, (uint4*)dense0);
                        }
                        count_launch();
                    }
                }
                ProfScope acc_scope(PROF_MSM_ACCUMULATE, stream);
                const uint32_t* off_in = bs;
                uint32_t* off_bufs[2] = {off_a, off_b};
                uint32_t* dense_bufs[2] = {dense_a, dense_b};
                const uint32_t* dense_in = records ? dense0 : nullptr;
                size_t bound = entries;                                  // upper bound on the level's input count
                for (int l = 0; l < levels; l++) {
                    uint32_t* off_out = off_bufs[l & 1];
                    uint32_t* dense_out = dense_bufs[l & 1];
                    k_halve_counts<<<(tb + 256) / 256, 256, 0, stream>>>(off_in, cnt_tmp, tb);
                    CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, cnt_tmp, off_out, (int)(tb + 1), stream));
                    bound = bound / 2 + tb;                              // Σ ceil(cnt/2) ≤ Σ cnt/2 + #buckets
                    // Whole waves: 148 SMs × 4 resident CTAs × 128 threads = 75776 lanes run at once; give every lane
                    // the same number T of outputs and launch an integer number of such waves, so no partial last wave
                    // idles most of the machine (a level is one long-running CTA per slot, not many short ones).
                    const size_t wave = (size_t)sm_count * (pair_v1 ? 4 : pair_minb) * 128;
                    size_t waves = (bound + 1024 * wave - 1) / (1024 * wave);
                    if (pair_waves) waves = pair_waves;
                    size_t T = (bound + waves * wave - 1) / (waves * wave);
                    const size_t nthreads = (bound + T - 1) / T;
                    const unsigned lgrid = (unsigned)((nthreads + 127) / 128);
                    // level 0 reads absolute positions of `sorted` (off_in = bs); its outputs and all later levels are
                    // group-relative (the scans start at 0)
                    if (pair_v1) {
                        if (l == 0 && !records)
                            k_pair_level<true><<<lgrid, 128, 0, stream>>>(gather_src, sorted, nullptr, off_in, off_out, tb, (uint32_t)T, prefix, dense_out);
                        else
                            k_pair_level<false><<<lgrid, 128, 0, stream>>>(nullptr, nullptr, dense_in, off_in, off_out, tb, (uint32_t)T, prefix, dense_out);
                    } else {
                        const unsigned dgrid = (unsigned)((bound + 255) / 256);
                        if (l == 0 && !records) {
                            k_pair_desc<true><<<dgrid, 256, 0, stream>>>(sorted, off_in, off_out, tb, desc, nullptr);
                            if (pair_minb == 3) k_pair_level2<true, 3><<<lgrid, 128, PAIR2_SMEM, stream>>>(gather_src, desc, off_out + tb, (uint32_t)T, prefix, dense_out, sm_slots);
                            else k_pair_level2<true, 4><<<lgrid, 128, PAIR2_SMEM, stream>>>(gather_src, desc, off_out + tb, (uint32_t)T, prefix, dense_out, sm_slots);
                        } else {
                            k_pair_desc<false><<<dgrid, 256, 0, stream>>>(nullptr, off_in, off_out, tb, desc, l == 0 ? bs : nullptr);
                            if (pair_minb == 3) k_pair_level2<false, 3><<<lgrid, 128, PAIR2_SMEM, stream>>>(dense_in, desc, off_out + tb, (uint32_t)T, prefix, dense_out, sm_slots);
                            else k_pair_level2<false, 4><<<lgrid, 128, PAIR2_SMEM, stream>>>(dense_in, desc, off_out + tb, (uint32_t)T, prefix, dense_out, sm_slots);
                        }
                        count_launch(1);
                    }
                    count_launch(3);
                    off_in = off_out;
                    dense_in = dense_out;
                }
                k_items_from_offsets<<<(tb + 256) / 256, 256, 0, stream>>>(off_in, items, tb, plan.cap, large_hot ? hot_dev : nullptr, (uint32_t)hot_max, hot_keep);
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, items, item_start, (int)(tb + 1), stream));
                const size_t group_items = (size_t)tb + bound / plan.cap + 1;
                items_bound = ((set_cap >> levels) + 1) / plan.cap + 1;
                items_launched = group_items;
                k_bucket_accumulate_dense<<<(unsigned)((group_items + MSM_ACC_THREADS - 1) / MSM_ACC_THREADS), MSM_ACC_THREADS, 0, stream>>>(
                    dense_in, off_in, item_start, tb, plan.cap, partial);
                count_launch(3);
            }
            ProfScope red_scope(PROF_MSM_REDUCE, stream);
            int quad_rounds = 0;
            if (quad_fold) {
                // the quad reduction adds up to 32 partials per bucket itself; buckets with more are folded 32:1 per round from the
                // device's hot list, as many rounds as the worst case needs
                size_t worst = acc_q8 ? 1 : items_bound;
                const uint32_t* p_in = partial; uint32_t* p_out = partial2;
                while (worst > 32) {
                    k_fold_hot_quad<<<(unsigned)sm_count * 8, 128, 0, stream>>>(p_in, item_start, hot_dev, (uint32_t)hot_max, (uint32_t)quad_rounds, 32u, p_out);
                    count_launch();
                    worst = (worst + 31) / 32;
                    quad_rounds++;
                    const uint32_t* t1 = p_in; p_in = p_out; p_out = (uint32_t*)t1;
                }
                final_partial = partial; final_start = item_start;
            } else if (large_hot) {
                // buckets with more than one item partial (the hot ones: few) are folded to one by the list-driven quad kernel
                size_t worst = items_bound;
                const uint32_t* p_in = partial; uint32_t* p_out = partial2;
                while (worst > 1) {
                    k_fold_hot_quad<<<(unsigned)sm_count * 8, 128, 0, stream>>>(p_in, item_start, hot_dev, (uint32_t)hot_max, (uint32_t)quad_rounds, 1u, p_out);
                    count_launch();
                    worst = (worst + 31) / 32;
                    quad_rounds++;
                    const uint32_t* t1 = p_in; p_in = p_out; p_out = (uint32_t*)t1;
                }
                final_partial = partial; final_start = item_start;
            } else {
                // fold item partials 32:1 until no bucket can hold more than one (worst case: all entries in one bucket).
                // `worst` (≥ the item count of any single bucket) only decides when to stop; the launch covers
                // Σ_b ceil(items_b / 32) ≤ #buckets + total/32 outputs, where `total_bound` bounds the items of ALL buckets
                // (a few hot buckets over a full background need far more outputs than tb + worst/32).
                size_t worst = items_bound;
                size_t total_bound = items_launched;
                uint32_t* p_in = partial; uint32_t* p_out = partial2;
                uint32_t* st_in = item_start; uint32_t* st_out = items2;
                while (worst > 1) {
                    k_group_counts<<<(tb + 256) / 256, 256, 0, stream>>>(st_in, cnt_tmp, tb);
                    CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, cnt_tmp, st_out, (int)(tb + 1), stream));
                    const size_t out_bound = (size_t)tb + total_bound / 32 + 1;
                    total_bound = out_bound;
                    k_partial_group_sum<<<(unsigned)((out_bound + 127) / 128), 128, 0, stream>>>(p_in, st_in, st_out, tb, p_out);
                    count_launch(3);
                    worst = (worst + 31) / 32;
                    uint32_t* t1 = p_in; p_in = p_out; p_out = t1;
                    uint32_t* t2 = st_in; st_in = st_out; st_out = t2;
                }
                final_partial = p_in; final_start = st_in;
            }
            uint32_t* group_sums = d_window_sums + (size_t)w0 * XYZZ_WORDS;
            if (warp_reduce && quad_path != 0) {
                const uint32_t qchunks = (plan.nbuckets + 7u) / 8u;                     // ≤ 128
                k_bucket_reduce_quad<<<(wn * qchunks * 32u + 127u) / 128u, 128, 0, stream>>>(final_partial, quad_fold ? partial2 : final_partial, final_start,
                                                                                              quad_rounds, plan.nbuckets, qchunks, wn, red_a);
                const uint32_t* ent = red_a;
                uint32_t m = qchunks;
                int lgw = 3;
                if (m > 64u) {                                                          // c = 11: 128 entries → 16
                    const uint32_t groups = (m + 7u) / 8u;
                    k_combine_level_quad<<<(wn * groups * 32u + 127u) / 128u, 128, 0, stream>>>(ent, m, wn, lgw, red_b);
                    count_launch();
                    ent = red_b; m = groups; lgw += 3;
                }
                k_window_combine_quad<<<wn, 32u * ((m + 7u) / 8u), 0, stream>>>(ent, m, lgw, group_sums);
                count_launch(2);
                continue;
            }
            if (warp_reduce) {
                const uint32_t wchunks = (plan.nbuckets + 31u) / 32u;
                k_bucket_reduce_warp<<<(wn * wchunks * 32u + 127u) / 128u, 128, 0, stream>>>(final_partial, final_start, plan.nbuckets, wchunks, wn, red_a);
                k_window_combine_warp<<<wn, 32, 0, stream>>>(red_a, wchunks, group_sums);
                count_launch(2);
                continue;
            }
            const uint32_t nthreads = chunks_per_set * wn;
            if (quad_path != 0 && quad_tail_large) {
                // one thread per 16-bucket chunk leaves its (weighted sum, sum) pair; the chunk offsets are applied by 8:1 quad-lane
                // folds (span 16 → 128 → 1024 → …), the last ≤ 64 entries of a set inside one CTA
                k_bucket_reduce<true><<<(nthreads + 127) / 128, 128, 0, stream>>>(final_partial, final_start, plan.nbuckets, chunk, chunks_per_set, wn, red_a,
                                                                                    partial2, large_hot ? quad_rounds : 0);
                count_launch(1);
                const uint32_t* ent = red_a;
                uint32_t* other = red_b;
                uint32_t m = chunks_per_set;
                int lgw = 0;
                while ((1u << lgw) < chunk) lgw++;
                while (m > 64u) {
                    const uint32_t groups = (m + 7u) / 8u;
                    if ((size_t)wn * groups >= 2048)                    // many groups: one quad each, sequential
                        k_combine_level_quadseq<<<(unsigned)(((size_t)wn * groups * 4 + 127) / 128), 128, 0, stream>>>(ent, m, wn, lgw, other);
                    else
                        k_combine_level_quad<<<(wn * groups * 32u + 127u) / 128u, 128, 0, stream>>>(ent, m, wn, lgw, other);
                    count_launch();
                    uint32_t* done = other; other = (uint32_t*)ent; ent = done;
                    m = groups; lgw += 3;
                }
                k_window_combine_quad<<<wn, 32u * ((m + 7u) / 8u), 0, stream>>>(ent, m, lgw, group_sums);
                count_launch();
                continue;
            }
            k_bucket_reduce<false><<<(nthreads + 127) / 128, 128, 0, stream>>>(final_partial, final_start, plan.nbuckets, chunk, chunks_per_set, wn, red_a);
            count_launch(1);
            // tree over the per-chunk sums: groups of `tree` until one point per bucket set remains
            uint32_t per_row = chunks_per_set;
            const uint32_t* src = red_a;
            uint32_t* bufs[2] = {red_b, red_a};
            int which = 0;
            while (per_row > 1) {
                uint32_t out_per_row = (per_row + tree - 1) / tree;
                uint32_t* target = out_per_row == 1 ? group_sums : bufs[which];
                uint32_t nt = out_per_row * wn;
                k_group_sum<<<(nt + 127) / 128, 128, 0, stream>>>(src, per_row, tree, out_per_row, wn, target);
                count_launch();
                src = target; which ^= 1; per_row = out_per_row;
            }
            if (chunks_per_set == 1)
                CUDA_TRY(cudaMemcpyAsync(group_sums, red_a, (size_t)wn * XYZZ_WORDS * 4, cudaMemcpyDeviceToDevice, stream));
        }
        CUDA_TRY(cudaGetLastError());
    }
done:
    scratch_release(block, scratch_bytes, stream, ds);
    return rc;
}

// ---- building blocks shared with the G2 path (msm_g2.cu): the curve-independent half of Pippenger ----
size_t msm_scan_bytes(size_t count) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)count);
    return bytes < 16 ? 16 : bytes;
}
int msm_exclusive_scan(void* tmp, size_t tmp_bytes, const uint32_t* in, uint32_t* out, size_t count, cudaStream_t stream) {
    count_launch(2);
    return (int)cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)count, stream);
}
int msm_sort_indices(const MsmPlan& plan, const void* d_scalars, size_t n, int mont, uint32_t* hist, uint32_t* bucket_start, uint32_t* cursors,
                     uint32_t* sorted, void* cub_tmp, size_t cub_bytes, uint32_t* d_flags, cudaStream_t stream) {
    const uint32_t TB = (uint32_t)plan.nwin * plan.nbuckets;
    int rc = (int)cudaMemsetAsync(hist, 0, (size_t)(TB + 1) * 4, stream);
    if (rc) return rc;
    const unsigned grid = (unsigned)((n + 255) / 256);
    const uint32_t* sc = (const uint32_t*)d_scalars;
    if (mont) k_digits<false, true><<<grid, 256, 0, stream>>>(sc, n, plan.c, plan.nwin, plan.nbuckets, hist, nullptr, 0, 0u, 0u, d_flags);
    else k_digits<false, false><<<grid, 256, 0, stream>>>(sc, n, plan.c, plan.nwin, plan.nbuckets, hist, nullptr, 0, 0u, 0u, d_flags);
    if ((rc = msm_exclusive_scan(cub_tmp, cub_bytes, hist, bucket_start, (size_t)TB + 1, stream)) != 0) return rc;
    if ((rc = (int)cudaMemcpyAsync(cursors, bucket_start, (size_t)(TB + 1) * 4, cudaMemcpyDeviceToDevice, stream)) != 0) return rc;
    if (mont) k_digits<true, true><<<grid, 256, 0, stream>>>(sc, n, plan.c, plan.nwin, plan.nbuckets, cursors, sorted, 0, 0u, 0u, d_flags);
    else k_digits<true, false><<<grid, 256, 0, stream>>>(sc, n, plan.c, plan.nwin, plan.nbuckets, cursors, sorted, 0, 0u, 0u, d_flags);
    count_launch(3);
    return (int)cudaGetLastError();
}
int msm_items_per_bucket(const uint32_t* hist, uint32_t* items, uint32_t total_buckets, uint32_t cap, cudaStream_t stream) {
    k_items_per_bucket<<<(total_buckets + 256) / 256, 256, 0, stream>>>(hist, items, total_buckets, cap, nullptr, 0u, 32u);
    count_launch();
    return (int)cudaGetLastError();
}
int msm_group_counts(const uint32_t* start_in, uint32_t* cnt_out, uint32_t total_buckets, cudaStream_t stream) {
    k_group_counts<<<(total_buckets + 256) / 256, 256, 0, stream>>>(start_in, cnt_out, total_buckets);
    count_launch();
    return (int)cudaGetLastError();
}

int msm_window_sums_device(uint32_t* d_window_sums, uint32_t* d_flags, const MsmPlan& plan, const void* d_points, size_t stride,
                           const void* d_scalars, size_t npoints, cudaStream_t stream) {
    MsmBases b{d_points, stride, npoints};
    MsmSegment sg{d_scalars, npoints, 0u, 0u, 0};
    return msm_core(d_window_sums, d_flags, plan, &b, 1, nullptr, 0, &sg, 1, 1, stream);
}

MsmPlan msm_make_plan_precomputed(size_t npoints) {
    MsmPlan p;
    int lg = ceil_log2(npoints < 2 ? 2 : npoints);
    // one bucket set for all windows ⇒ the bucket reduction is paid once and wide windows are cheap.  Re-swept in round 2 with
    // the record scatter feeding dense pair levels (tools/tune_precomputed.py, profiles/r2j_tune_pre.log): c = 20 from 2^22
    // (13 windows; 80.8 ms at 2^24 against 90.9 ms without tables), c = 17 at 2^20–2^21.
    int c = lg >= 22 ? 20 : lg >= 20 ? 17 : lg >= 8 ? lg - 2 : 6;
    if (const char* e = getenv("SNARKVM_B200_MSM_PRE_C")) { int v = atoi(e); if (v >= 2 && v <= 24) c = v; }
    p.c = c;
    p.nwin = 253 / c + 1;
    p.nbuckets = 1u << (c - 1);
    size_t total = npoints * (size_t)p.nwin;
    size_t cap = total / 300000 + 1;
    if (cap < 16) cap = 16;
    p.cap = (uint32_t)cap;
    int levels = lg >= 24 ? 5 : lg >= 22 ? 4 : lg >= 20 ? 3 : 1;
    if (const char* e = getenv("SNARKVM_B200_MSM_PRE_LEVELS")) { int v = atoi(e); if (v >= 1 && v <= 16) levels = v; }
    p.levels = levels;
    return p;
}

int msm_precompute_tables_device(uint32_t* d_table, const MsmPlan& plan, const void* d_points, size_t stride, size_t npoints,
                                 cudaStream_t stream) {
    if (stride < 104 || (stride & 7) || npoints == 0) return (int)cudaErrorInvalidValue;
    k_precompute_tables<<<(unsigned)((npoints + 127) / 128), 128, 0, stream>>>((const uint8_t*)d_points, stride, npoints, plan.c, plan.nwin, d_table);
    count_launch();
    return (int)cudaGetLastError();
}

int msm_precomputed_sum_device(uint32_t* d_sum, uint32_t* d_flags, const MsmPlan& plan, const uint32_t* d_table, size_t table_n, const void* d_scalars,
                               size_t nscalars, int mont, cudaStream_t stream) {
    MsmSegment sg{d_scalars, nscalars, 0u, 0u, mont};
    return msm_core(d_sum, d_flags, plan, nullptr, 0, d_table, table_n, &sg, 1, 1, stream);
}

// ---------------------------------------------------------------------------
// Synthetic bases: P_i = h(seed, i)·G
// ---------------------------------------------------------------------------
FF_DEV uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__constant__ uint32_t G1_GEN_X[12] = {0xec301b95u, 0x1042a645u, 0xc1060f28u, 0x5a990780u, 0xa9007a5bu, 0x684a8ab3u,
                                      0x257ba63fu, 0x1c35a184u, 0xfea8e32eu, 0xb2b2abd2u, 0x23fb2017u, 0x017df3a2u};   // g1.rs:225-236 (Montgomery)
__constant__ uint32_t G1_GEN_Y[12] = {0xe2801ab9u, 0xbc5a1ae8u, 0xcbfe13b0u, 0xd4f3c861u, 0x4e949f13u, 0xecdd5ffcu,
                                      0x7503667du, 0x8f87199bu, 0x7dc4fe1cu, 0x0f0b1b83u, 0x053eaabeu, 0x004bcc7eu};   // g1.rs:242-253

__global__ void __launch_bounds__(128) k_generate_bases(uint8_t* points, size_t n, size_t stride, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = splitmix64(seed ^ splitmix64((uint64_t)i));
    if (k == 0) k = 1;
    AffinePoint g;
#pragma unroll
    for (int j = 0; j < 12; j++) { g.x.v[j] = G1_GEN_X[j]; g.y.v[j] = G1_GEN_Y[j]; }
    g.inf = false;
    XYZZ acc = XYZZ::infinity();
    bool started = false;
    for (int b = 63; b >= 0; b--) {
        if (started) acc.dbl();
        if ((k >> b) & 1ull) { acc.add_affine(g, false); started = true; }
    }
    store_affine(points, stride, i, acc.to_affine());
}

// ---------------------------------------------------------------------------
// SRS ingest (first half): uncompressed canonical points of a `.usrs` file
// (x LE 48 B, y LE 48 B, flags in the top two bits of the last byte: bit 6 = infinity —
// utilities/src/serialize/flags.rs:72-98, curves/src/templates/macros.rs:86-96,136) → the reference's
// in-memory Affine image (Montgomery x, y, infinity flag) straight in HBM, with the on-curve check
// y² = x³ + 1 (affine.rs:205-215) and the canonical-range check (< q) counted into *invalid.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_srs_decode(const uint8_t* __restrict__ in, size_t n, uint8_t* __restrict__ out, size_t stride,
                                                     uint32_t* __restrict__ invalid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in + i * 96);     // 96-byte records are 4-byte aligned after the 8-byte header
    Fq x, y;
#pragma unroll
    for (int k = 0; k < 12; k++) { x.v[k] = __ldg(w + k); y.v[k] = __ldg(w + 12 + k); }
    const uint32_t flags = y.v[11] >> 30;                                   // bit 31 = y sign (unused when uncompressed), bit 30 = infinity
    y.v[11] &= 0x3fffffffu;
    AffinePoint a;
    bool bad = false;
    if (flags & 1u) {
        a.x = Fq::zero(); a.y = Fq::one(); a.inf = true;                    // Affine::zero(), affine.rs:57-59
        bad = (flags & 2u) != 0u;                                           // (sign, infinity) both set is not a valid encoding
    } else {
        // canonical range: x, y < q  ⇔  (v − q) borrows
        Fq tx = x, ty = y;
        tx.final_sub(); ty.final_sub();
        bad = (tx != x) || (ty != y);
        a.x = x.to_mont(); a.y = y.to_mont(); a.inf = false;
        Fq lhs = a.y.sqr(), rhs = a.x.sqr() * a.x + Fq::one();
        bad = bad || (lhs != rhs);
    }
    if (bad) atomicAdd(invalid, 1u);
    store_affine(out, stride, i, a);
}
int srs_decode_device(void* d_out, size_t stride, const void* d_in, size_t npoints, uint32_t* d_invalid, cudaStream_t stream) {
    if (stride < 104 || (stride & 7) || ((uintptr_t)d_in & 3)) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(d_invalid, 0, 4, stream);
    if (e != cudaSuccess) return (int)e;
    if (npoints == 0) return 0;
    k_srs_decode<<<(unsigned)((npoints + 127) / 128), 128, 0, stream>>>((const uint8_t*)d_in, npoints, (uint8_t*)d_out, stride, d_invalid);
    count_launch();
    return (int)cudaGetLastError();
}

// Self-test of the warp-cooperative field arithmetic (ff.cuh coop_mul / coop_inverse) against the per-thread multiplier:
// warp w takes pseudo-random elements a, b (and the edge values 0, 1, p − 1), checks coop_mul(a, b) == a·b limb by limb and
// coop_inverse(a)·a == 1.  *mismatches counts failures.
__global__ void __launch_bounds__(128) k_selftest_coop(uint32_t nwarps, uint64_t seed, uint32_t* __restrict__ mismatches) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    Fq a, b;
#pragma unroll
    for (int k = 0; k < 12; k += 2) {
        uint64_t ra = splitmix64(seed + 0x100ull * w + (uint64_t)k), rb = splitmix64(~seed + 0x100ull * w + (uint64_t)k);
        a.v[k] = (uint32_t)ra; a.v[k + 1] = (uint32_t)(ra >> 32); b.v[k] = (uint32_t)rb; b.v[k + 1] = (uint32_t)(rb >> 32);
    }
    a.v[11] &= 0x00ffffffu; b.v[11] &= 0x00ffffffu;             // < 2^376 < q: already reduced
    if (w == 0) a = Fq::zero();
    if (w == 1) a = Fq::one();
    if (w == 2) a = Fq::zero() - Fq::one();
    if (w == 3) { a = Fq::zero() - Fq::one(); b = a; }
    if (w == 4) {                                              // long carry chains: 2^352 − 1 times R
#pragma unroll
        for (int k = 0; k < 12; k++) a.v[k] = k < 11 ? 0xffffffffu : 0u;
        b = Fq::one();
    }
    const uint32_t p = coop_mod_limb<FqParams>(lane);
    uint32_t al = 0u, bl = 0u;
#pragma unroll
    for (int k = 0; k < 12; k++) { if (lane == k) { al = a.v[k]; bl = b.v[k]; } }
    const uint32_t prod = coop_mul<FqParams>(al, bl, p, lane);
    const Fq want = a * b;
    uint32_t wl = 0u;
#pragma unroll
    for (int k = 0; k < 12; k++) if (lane == k) wl = want.v[k];
    bool bad = prod != wl;
    const Fq inv = coop_inverse<FqParams>(a);
    const Fq one = inv * a;
    if (a.is_zero()) bad = bad || !inv.is_zero();
    else bad = bad || (one != Fq::one()) || (inv != a.inverse());
    // the Karatsuba multiplier against the word-serial schoolbook one, per lane, Fq and Fr: random operands, operands whose
    // halves are equal / swapped (every sign combination of the middle term), 0, 1, p − 1
    {
        const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
        Fq x, y;
        Fr u, v;
#pragma unroll
        for (int k = 0; k < 12; k += 2) {
            uint64_t rx = splitmix64(seed * 3 + 0x1000ull * t + (uint64_t)k), ry = splitmix64(~seed * 5 + 0x1000ull * t + (uint64_t)k);
            x.v[k] = (uint32_t)rx; x.v[k + 1] = (uint32_t)(rx >> 32); y.v[k] = (uint32_t)ry; y.v[k + 1] = (uint32_t)(ry >> 32);
        }
        x.v[11] &= 0x00ffffffu; y.v[11] &= 0x00ffffffu;
        if ((t & 7u) == 1u) { for (int k = 0; k < 6; k++) { y.v[k] = x.v[k + 6]; y.v[k + 6] = x.v[k]; } y.v[11] &= 0x00ffffffu; y.v[5] = x.v[11]; }
        if ((t & 7u) == 2u) { for (int k = 0; k < 6; k++) x.v[k + 6] = x.v[k]; x.v[11] &= 0x00ffffffu; }
        if (t == 8u) x = Fq::zero();
        if (t == 9u) y = Fq::one();
        if (t == 10u) { x = Fq::zero() - Fq::one(); y = x; }
        if (t == 11u) { x = Fq::zero() - Fq::one(); }
#pragma unroll
        for (int k = 0; k < 8; k++) { u.v[k] = x.v[k] ^ y.v[11 - k]; v.v[k] = y.v[k] + x.v[11 - k]; }
        u.v[7] &= 0x0fffffffu; v.v[7] &= 0x0fffffffu;
        if ((t & 7u) == 3u) { for (int k = 0; k < 4; k++) v.v[k + 4] = v.v[k]; v.v[7] &= 0x0fffffffu; }
        if (t == 12u) { u = Fr::zero() - Fr::one(); v = u; }
        bool bad2 = Fq::mul_karatsuba(x, y) != Fq::mul_inline(x, y) || Fq::mul_karatsuba(x, x) != Fq::sqr_inline(x);
        bad2 = bad2 || Fr::mul_karatsuba(u, v) != Fr::mul_inline(u, v) || Fr::mul_karatsuba(v, v) != Fr::sqr_inline(v);
        bad = bad || bad2;
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(mismatches, 1u);
}
int selftest_coop_device(uint32_t nwarps, uint64_t seed, uint32_t* d_mismatches, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_mismatches, 0, 4, stream);
    if (e != cudaSuccess) return (int)e;
    k_selftest_coop<<<(nwarps * 32 + 127) / 128, 128, 0, stream>>>(nwarps, seed, d_mismatches);
    count_launch();
    return (int)cudaGetLastError();
}

// P_i = s_i·G for canonical 253-bit scalars s_i in HBM (a universal setup with a KNOWN trapdoor for tests and benches:
// s_i = β^i gives powers_of_beta_g, s_i = γβ^i powers_of_beta_times_gamma_g; kzg10/data_structures.rs UniversalParams)
__global__ void __launch_bounds__(128) k_generator_mul(uint8_t* points, size_t n, size_t stride, const uint32_t* __restrict__ scalars) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = scalars[8 * i + j];
    AffinePoint g;
#pragma unroll
    for (int j = 0; j < 12; j++) { g.x.v[j] = G1_GEN_X[j]; g.y.v[j] = G1_GEN_Y[j]; }
    g.inf = false;
    XYZZ acc = XYZZ::infinity();
    bool started = false;
#pragma unroll 1
    for (int w = 7; w >= 0; w--) {
        uint32_t kw = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) if (j == w) kw = k[j];
#pragma unroll 1
        for (int b = 31; b >= 0; b--) {
            if (started) acc.dbl();
            if ((kw >> b) & 1u) { acc.add_affine(g, false); started = true; }
        }
    }
    store_affine(points, stride, i, acc.to_affine());
}
int msm_generator_mul_device(void* d_points, size_t npoints, size_t stride, const void* d_scalars, cudaStream_t stream) {
    if (stride < 104 || (stride & 7)) return (int)cudaErrorInvalidValue;
    if (npoints == 0) return 0;
    k_generator_mul<<<(unsigned)((npoints + 127) / 128), 128, 0, stream>>>((uint8_t*)d_points, npoints, stride, (const uint32_t*)d_scalars);
    count_launch();
    return (int)cudaGetLastError();
}

int msm_generate_bases_device(void* d_points, size_t npoints, size_t stride, uint64_t seed, cudaStream_t stream) {
    if (stride < 104 || (stride & 7)) return (int)cudaErrorInvalidValue;
    if (npoints == 0) return 0;
    k_generate_bases<<<(unsigned)((npoints + 127) / 128), 128, 0, stream>>>((uint8_t*)d_points, npoints, stride, seed);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace b200
