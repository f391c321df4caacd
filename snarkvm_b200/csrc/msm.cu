// See msm.cuh for the map from reference functions to kernels.
#include "msm.cuh"

#include <atomic>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_scan.cuh>

#define FF_CALL_MUL 1
#include "ec.cuh"

namespace b200 {

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n); }

void ensure_pool_configured() {
    static std::once_flag once[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::call_once(once[dev & 63], [dev] {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    });
}

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

MsmPlan msm_make_plan(size_t npoints) {
    MsmPlan p;
    int lg = ceil_log2(npoints < 2 ? 2 : npoints);
    // Window bits from a sweep on B200 (tools/tune_msm.py, profiles/tune_msm_r1.log): wider windows mean fewer
    // bucket additions (n·W) but more buckets to reduce and shorter, more divergent bucket runs.
    int c = lg <= 8 ? 4 : lg <= 12 ? lg - 4 : lg <= 18 ? 11 : lg == 19 ? 16 : lg == 20 ? 15 : lg <= 22 ? 16 : 17;
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
    int levels = lg >= 21 ? 4 : lg == 20 ? 1 : 0;
    while (levels > 0 && ((npoints >> (c - 1)) >> levels) < 2) levels--;
    if (const char* e = getenv("SNARKVM_B200_MSM_LEVELS")) { int v = atoi(e); if (v >= 0 && v <= 16) levels = v; }
    p.levels = levels;
    return p;
}

// ---------------------------------------------------------------------------
// Signed-digit recoding of a canonical 253-bit scalar (8 little-endian u32 words).
// digit_w ∈ [-2^(c-1), 2^(c-1)]; returns magnitude and sign for window w given the
// running carry (sequential over w).
// ---------------------------------------------------------------------------
// flat != 0 (precomputed tables 2^{c·w}·P_i): every window feeds the SAME bucket set and the entry names record
// w·n + i of the table instead of point i.
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(const uint32_t* __restrict__ scalars, size_t n, int c, int nwin,
                                                uint32_t nbuckets, uint32_t* __restrict__ counters /* hist or cursors */,
                                                uint32_t* __restrict__ sorted, size_t flat /* 0, or the table's points per window */) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    {
        const uint4* q = reinterpret_cast<const uint4*>(scalars + 8 * i);
        uint4 a = __ldg(q), b = __ldg(q + 1);
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    }
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
            uint32_t slot = (flat ? 0u : (uint32_t)w * nbuckets) + (mag - 1u);
            if (SCATTER) {
                uint32_t pos = atomicAdd(&counters[slot], 1u);
                sorted[pos] = (uint32_t)(flat ? (size_t)w * flat + i : i) | (neg << 31);
            } else {
                atomicAdd(&counters[slot], 1u);
            }
        }
    }
}

__global__ void k_items_per_bucket(const uint32_t* __restrict__ hist, uint32_t* __restrict__ items, uint32_t total_buckets, uint32_t cap) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_buckets) items[i] = (hist[i] + cap - 1u) / cap;
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
__global__ void __launch_bounds__(MSM_ACC_THREADS, MSM_ACC_MINBLOCKS) k_bucket_accumulate(const uint8_t* __restrict__ points, size_t stride,
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
    AffinePoint p = load_affine(points, stride, e & 0x7fffffffu);
    for (uint32_t k = s0; k < s1; k++) {
        uint32_t e_cur = e;
        AffinePoint p_cur = p;
        if (k + 1 < s1) { e = sorted[k + 1]; p = load_affine(points, stride, e & 0x7fffffffu); }
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
// One thread per point walks the doubling chain in XYZZ and normalises every multiple (Fermat inversion each:
// a one-off cost per SRS, ≈ 9k Fq mul per point).
__global__ void __launch_bounds__(128) k_precompute_tables(const uint8_t* __restrict__ points, size_t stride, size_t n, int c, int nwin,
                                                            uint32_t* __restrict__ table) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    AffinePoint a = load_affine(points, stride, i);
    DensePoint d; d.x = a.x; d.y = a.y; d.inf = a.inf;
    store_dense(table + i * BASE_WORDS, d);
    XYZZ q = XYZZ::from_affine(a);
    for (int w = 1; w < nwin; w++) {
        for (int k = 0; k < c; k++) q.dbl();
        AffinePoint t = q.to_affine();
        d.x = t.x; d.y = t.y; d.inf = t.inf;
        store_dense(table + ((size_t)w * n + i) * BASE_WORDS, d);
    }
}

static constexpr int PAIR_THREADS = 128;
FF_DEV Fq shfl_up_fq(const Fq& a, int d) {
    Fq r;
#pragma unroll
    for (int j = 0; j < 12; j++) r.v[j] = __shfl_up_sync(0xffffffffu, a.v[j], d);
    return r;
}
FF_DEV Fq shfl_down_fq(const Fq& a, int d) {
    Fq r;
#pragma unroll
    for (int j = 0; j < 12; j++) r.v[j] = __shfl_down_sync(0xffffffffu, a.v[j], d);
    return r;
}
FF_DEV Fq shfl_idx_fq(const Fq& a, int l) {
    Fq r;
#pragma unroll
    for (int j = 0; j < 12; j++) r.v[j] = __shfl_sync(0xffffffffu, a.v[j], l);
    return r;
}
// One Fermat inversion per CTA instead of one per warp-lane: the 128 running products (all non-zero) go through shared memory,
// ONE warp multiplies them together (4 per lane, then a shuffle scan across lanes), inverts the total and unwinds.  In SIMT
// time an inversion costs a warp ≈ 515 Fq mul whether 1 or 32 lanes need it, so the per-thread version spends 4 × 515 per
// CTA and this one ≈ 540.  The inverting warp rotates with blockIdx so co-resident CTAs load different SM sub-partitions.
__device__ __noinline__ Fq cta_shared_inverse(const Fq& run, uint32_t* sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    run.store(sh + tid * 12);
    __syncthreads();
    if (warp == (int)(blockIdx.x & 3u)) {
        uint32_t* mine = sh + lane * 48;
        Fq a0 = Fq::load(mine), a1 = Fq::load(mine + 12), a2 = Fq::load(mine + 24), a3 = Fq::load(mine + 36);
        Fq p1 = a0 * a1, p2 = p1 * a2, p3 = p2 * a3;
        Fq incl = p3, suff = p3;                           // inclusive prefix / suffix products over lanes
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            Fq up = shfl_up_fq(incl, d), dn = shfl_down_fq(suff, d);
            if (lane >= d) incl = incl * up;
            if (lane + d < 32) suff = suff * dn;
        }
        Fq tinv = shfl_idx_fq(incl, 31).inverse();
        Fq before = shfl_up_fq(incl, 1), after = shfl_down_fq(suff, 1);
        Fq ip3 = tinv;                                     // 1 / p3 of this lane = tinv · Π(other lanes)
        if (lane > 0) ip3 = ip3 * before;
        if (lane < 31) ip3 = ip3 * after;
        Fq ip2 = ip3 * a3, ip1 = ip2 * a2;
        (ip3 * p2).store(mine + 36);                       // 1/a3
        (ip2 * p1).store(mine + 24);                       // 1/a2
        (ip1 * a0).store(mine + 12);                       // 1/a1
        (ip1 * a1).store(mine);                            // 1/a0
    }
    __syncthreads();
    return Fq::load(sh + tid * 12);
}

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

__global__ void k_halve_counts(const uint32_t* __restrict__ off_in, uint32_t* __restrict__ cnt_out, uint32_t total_buckets) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_buckets) cnt_out[i] = (off_in[i + 1] - off_in[i] + 1u) >> 1;
    else if (i == total_buckets) cnt_out[i] = 0;
}
__global__ void k_items_from_offsets(const uint32_t* __restrict__ off, uint32_t* __restrict__ items, uint32_t total_buckets, uint32_t cap) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_buckets) items[i] = (off[i + 1] - off[i] + cap - 1u) / cap;
    else if (i == total_buckets) items[i] = 0;
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
__global__ void __launch_bounds__(128) k_bucket_reduce(const uint32_t* __restrict__ partial, const uint32_t* __restrict__ item_start,
                                                        uint32_t nbuckets, uint32_t chunk, uint32_t chunks_per_window,
                                                        uint32_t nwin, uint32_t* __restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks_per_window * nwin) return;
    uint32_t w = t / chunks_per_window, j = t % chunks_per_window;
    uint32_t lo = j * chunk, hi = lo + chunk;                 // 0-based bucket indices [lo, hi)
    XYZZ running = XYZZ::infinity(), acc = XYZZ::infinity();
    for (uint32_t b = hi; b-- > lo;) {
        XYZZ s = bucket_sum(partial, item_start, w * nbuckets + b);
        running.add(s);
        acc.add(running);
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

// table != nullptr: "flat" mode over precomputed tables (record w·table_n + i = 2^{c·w}·P_i): all windows share one bucket
// set, so the pipeline sees ONE window of npoints·nwin entries and writes a single sum.
static int msm_core(uint32_t* d_window_sums, const MsmPlan& plan, const void* d_points, size_t stride, const uint32_t* table,
                    size_t table_n, const void* d_scalars, size_t npoints, cudaStream_t stream) {
    int rc = 0;
    ensure_pool_configured();
    const bool flat = table != nullptr;
    const uint32_t nwin_red = flat ? 1u : (uint32_t)plan.nwin;    // bucket sets to reduce
    const uint32_t TB = nwin_red * plan.nbuckets;                 // total buckets
    const size_t max_entries = npoints * (size_t)plan.nwin;
    if (npoints == 0 || npoints >= (1ull << 31) || max_entries >= (1ull << 32)) return (int)cudaErrorInvalidValue;
    if (flat && (table_n * (size_t)plan.nwin >= (1ull << 31) || plan.levels < 1 || npoints > table_n)) return (int)cudaErrorInvalidValue;
    const int levels = plan.levels;
    const size_t bucket_cap = flat ? max_entries : npoints;       // most entries a single bucket can hold

    // Everything after the bucket sort runs per GROUP of whole windows, so the dense scratch of the pair levels
    // (96 B per point per window) stays inside a budget: 2^24 points → all 15 windows at once (24 GB),
    // 2^26 points → 3 groups of 6/6/3 windows.
    size_t budget = (size_t)40 << 30;
    if (const char* e = getenv("SNARKVM_B200_MSM_SCRATCH_GB")) { long v = atol(e); if (v >= 1) budget = (size_t)v << 30; }
    uint32_t gw = nwin_red;
    if (levels > 0 && !flat) {
        size_t per_window = npoints * (size_t)96 + 1;
        size_t fit = budget / per_window;
        if (fit < 1) fit = 1;
        if (fit < gw) gw = (uint32_t)fit;
    }
    const uint32_t TBg = gw * plan.nbuckets;                      // buckets of the largest group
    const size_t entries_g = flat ? max_entries : npoints * (size_t)gw;
    const size_t max_items = (size_t)TBg + entries_g / plan.cap + 1;

    uint32_t *hist = nullptr, *bucket_start = nullptr, *cursors = nullptr, *items = nullptr, *item_start = nullptr;
    uint32_t *sorted = nullptr, *partial = nullptr, *red_a = nullptr, *red_b = nullptr;
    uint32_t *off_a = nullptr, *off_b = nullptr, *dense_a = nullptr, *dense_b = nullptr, *prefix = nullptr, *dense_bases = nullptr;
    uint32_t *partial2 = nullptr, *items2 = nullptr;
    const size_t dense_cap_a = entries_g / 2 + TBg + 1, dense_cap_b = entries_g / 4 + 2 * (size_t)TBg + 1;
    size_t pair_waves = 0;                       // 0 = fewest whole waves with T ≤ 1024 outputs per thread
    if (const char* e = getenv("SNARKVM_B200_MSM_PAIR_WAVES")) { long v = atol(e); if (v >= 1) pair_waves = (size_t)v; }
    int sm_count = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); if (sm_count <= 0) sm_count = 148; }
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    // The reduction tail is a chain of dependent point additions per thread (≈ 16 µs each for a lone warp): short chunks
    // and a narrow (8:1) tree keep that chain short — what matters at 2^16–2^20 points, where the tail is 20–50 % of the call.
    const uint32_t chunk = plan.nbuckets < 16u ? plan.nbuckets : 16u;
    const uint32_t tree = 8;
    const uint32_t chunks_per_window = plan.nbuckets / chunk;

    CUDA_TRY(cudaMallocAsync(&hist, (size_t)(TB + 1) * 4, stream));
    CUDA_TRY(cudaMallocAsync(&bucket_start, (size_t)(TB + 1) * 4, stream));
    CUDA_TRY(cudaMallocAsync(&cursors, (size_t)(TB + 1) * 4, stream));
    CUDA_TRY(cudaMallocAsync(&items, (size_t)(TBg + 1) * 4, stream));
    CUDA_TRY(cudaMallocAsync(&item_start, (size_t)(TBg + 1) * 4, stream));
    CUDA_TRY(cudaMallocAsync(&sorted, max_entries * 4, stream));
    CUDA_TRY(cudaMallocAsync(&partial, max_items * XYZZ_WORDS * 4, stream));
    CUDA_TRY(cudaMallocAsync(&partial2, ((size_t)TBg + max_items / 32 + 2) * XYZZ_WORDS * 4, stream));
    CUDA_TRY(cudaMallocAsync(&items2, (size_t)(TBg + 1) * 4, stream));
    if (levels > 0) {
        CUDA_TRY(cudaMallocAsync(&off_a, (size_t)(TBg + 1) * 4, stream));
        CUDA_TRY(cudaMallocAsync(&off_b, (size_t)(TBg + 1) * 4, stream));
        CUDA_TRY(cudaMallocAsync(&dense_a, dense_cap_a * DENSE_WORDS * 4, stream));
        if (levels > 1) CUDA_TRY(cudaMallocAsync(&dense_b, dense_cap_b * DENSE_WORDS * 4, stream));
        CUDA_TRY(cudaMallocAsync(&prefix, dense_cap_a * 12 * 4, stream));
        if (!flat) CUDA_TRY(cudaMallocAsync(&dense_bases, npoints * (size_t)BASE_WORDS * 4, stream));
    }
    CUDA_TRY(cudaMallocAsync(&red_a, (size_t)gw * chunks_per_window * XYZZ_WORDS * 4, stream));
    CUDA_TRY(cudaMallocAsync(&red_b, (size_t)gw * (chunks_per_window / tree + 1) * XYZZ_WORDS * 4, stream));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, hist, bucket_start, (int)(TB + 1), stream));
    CUDA_TRY(cudaMallocAsync(&cub_tmp, cub_bytes ? cub_bytes : 16, stream));

    CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)(TB + 1) * 4, stream));
    {
        // ---- bucket sort of all windows: histogram → offsets → scatter ----
        const unsigned grid = (unsigned)((npoints + 255) / 256);
        {
            ProfScope sort_scope(PROF_MSM_SORT, stream);
            k_digits<false><<<grid, 256, 0, stream>>>((const uint32_t*)d_scalars, npoints, plan.c, plan.nwin, plan.nbuckets, hist, nullptr, flat ? table_n : 0);
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, hist, bucket_start, (int)(TB + 1), stream));
            CUDA_TRY(cudaMemcpyAsync(cursors, bucket_start, (size_t)(TB + 1) * 4, cudaMemcpyDeviceToDevice, stream));
            k_digits<true><<<grid, 256, 0, stream>>>((const uint32_t*)d_scalars, npoints, plan.c, plan.nwin, plan.nbuckets, cursors, sorted, flat ? table_n : 0);
            count_launch(4);
        }
        const uint32_t* gather_src = flat ? table : dense_bases;
        if (levels > 0 && !flat) {
            ProfScope acc_scope(PROF_MSM_ACCUMULATE, stream);
            k_densify_bases<<<(unsigned)((npoints + 255) / 256), 256, 0, stream>>>((const uint8_t*)d_points, stride, npoints, dense_bases);
            count_launch();
        }
        for (uint32_t w0 = 0; w0 < nwin_red; w0 += gw) {
            const uint32_t wn = nwin_red - w0 < gw ? nwin_red - w0 : gw;                           // windows (bucket sets) in this group
            const uint32_t tb = wn * plan.nbuckets;
            const uint32_t* bs = bucket_start + (size_t)w0 * plan.nbuckets;                       // tb + 1 absolute offsets into `sorted`
            const size_t entries = flat ? max_entries : npoints * (size_t)wn;
            size_t items_bound = 1;                                                                // ≥ item count of any single bucket
            const uint32_t* final_partial = nullptr;
            const uint32_t* final_start = nullptr;
            if (levels == 0) {
                CUDA_TRY(cudaMemsetAsync(items, 0, (size_t)(tb + 1) * 4, stream));
                k_items_per_bucket<<<(tb + 255) / 256, 256, 0, stream>>>(hist + (size_t)w0 * plan.nbuckets, items, tb, plan.cap);
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, items, item_start, (int)(tb + 1), stream));
                count_launch(3);
                const size_t group_items = (size_t)tb + entries / plan.cap + 1;
                items_bound = bucket_cap / plan.cap + 1;                // a bucket belongs to one window: ≤ n entries
                ProfScope acc_scope(PROF_MSM_ACCUMULATE, stream);
                k_bucket_accumulate<<<(unsigned)((group_items + MSM_ACC_THREADS - 1) / MSM_ACC_THREADS), MSM_ACC_THREADS, 0, stream>>>(
                    (const uint8_t*)d_points, stride, sorted, bs, item_start, tb, plan.cap, partial);
            } else {
                ProfScope acc_scope(PROF_MSM_ACCUMULATE, stream);
                const uint32_t* off_in = bs;
                uint32_t* off_bufs[2] = {off_a, off_b};
                uint32_t* dense_bufs[2] = {dense_a, dense_b};
                const uint32_t* dense_in = nullptr;
                size_t bound = entries;                                  // upper bound on the level's input count
                for (int l = 0; l < levels; l++) {
                    uint32_t* off_out = off_bufs[l & 1];
                    uint32_t* dense_out = dense_bufs[l & 1];
                    k_halve_counts<<<(tb + 256) / 256, 256, 0, stream>>>(off_in, cursors, tb);
                    CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, cursors, off_out, (int)(tb + 1), stream));
                    bound = bound / 2 + tb;                              // Σ ceil(cnt/2) ≤ Σ cnt/2 + #buckets
                    // Whole waves: 148 SMs × 4 resident CTAs × 128 threads = 75776 threads run at once; give every thread
                    // the same number T of outputs and launch an integer number of such waves, so no partial last wave
                    // idles most of the machine (a level is one long-running CTA per slot, not many short ones).
                    const size_t wave = (size_t)sm_count * 4 * 128;
                    size_t waves = (bound + 1024 * wave - 1) / (1024 * wave);
                    if (pair_waves) waves = pair_waves;
                    size_t T = (bound + waves * wave - 1) / (waves * wave);
                    const size_t nthreads = (bound + T - 1) / T;
                    const unsigned lgrid = (unsigned)((nthreads + 127) / 128);
                    // level 0 reads absolute positions of `sorted` (off_in = bs); its outputs and all later levels are
                    // group-relative (the scans start at 0)
                    if (l == 0)
                        k_pair_level<true><<<lgrid, 128, 0, stream>>>(gather_src, sorted, nullptr, off_in, off_out, tb, (uint32_t)T, prefix, dense_out);
                    else
                        k_pair_level<false><<<lgrid, 128, 0, stream>>>(nullptr, nullptr, dense_in, off_in, off_out, tb, (uint32_t)T, prefix, dense_out);
                    count_launch(4);
                    off_in = off_out;
                    dense_in = dense_out;
                }
                k_items_from_offsets<<<(tb + 256) / 256, 256, 0, stream>>>(off_in, items, tb, plan.cap);
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, items, item_start, (int)(tb + 1), stream));
                const size_t group_items = (size_t)tb + bound / plan.cap + 1;
                items_bound = ((bucket_cap >> levels) + 1) / plan.cap + 1;
                k_bucket_accumulate_dense<<<(unsigned)((group_items + MSM_ACC_THREADS - 1) / MSM_ACC_THREADS), MSM_ACC_THREADS, 0, stream>>>(
                    dense_in, off_in, item_start, tb, plan.cap, partial);
                count_launch(4);
            }
            ProfScope red_scope(PROF_MSM_REDUCE, stream);
            {
                // fold item partials 32:1 until no bucket can hold more than one (worst case: all entries in one bucket)
                size_t worst = items_bound;
                uint32_t* p_in = partial; uint32_t* p_out = partial2;
                uint32_t* st_in = item_start; uint32_t* st_out = items2;
                while (worst > 1) {
                    k_group_counts<<<(tb + 256) / 256, 256, 0, stream>>>(st_in, cursors, tb);
                    CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, cursors, st_out, (int)(tb + 1), stream));
                    const size_t out_bound = (size_t)tb + (worst + 31) / 32;
                    k_partial_group_sum<<<(unsigned)((out_bound + 127) / 128), 128, 0, stream>>>(p_in, st_in, st_out, tb, p_out);
                    count_launch(4);
                    worst = (worst + 31) / 32;
                    uint32_t* t1 = p_in; p_in = p_out; p_out = t1;
                    uint32_t* t2 = st_in; st_in = st_out; st_out = t2;
                }
                final_partial = p_in; final_start = st_in;
            }
            uint32_t* group_sums = d_window_sums + (size_t)w0 * XYZZ_WORDS;
            const uint32_t nthreads = chunks_per_window * wn;
            k_bucket_reduce<<<(nthreads + 127) / 128, 128, 0, stream>>>(final_partial, final_start, plan.nbuckets, chunk, chunks_per_window, wn, red_a);
            count_launch(1);
            // tree over the per-chunk sums: groups of `tree` until one point per window remains
            uint32_t per_row = chunks_per_window;
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
            if (chunks_per_window == 1)
                CUDA_TRY(cudaMemcpyAsync(group_sums, red_a, (size_t)wn * XYZZ_WORDS * 4, cudaMemcpyDeviceToDevice, stream));
        }
        CUDA_TRY(cudaGetLastError());
    }
done:
    cudaFreeAsync(hist, stream); cudaFreeAsync(bucket_start, stream); cudaFreeAsync(cursors, stream);
    cudaFreeAsync(items, stream); cudaFreeAsync(item_start, stream); cudaFreeAsync(sorted, stream);
    cudaFreeAsync(off_a, stream); cudaFreeAsync(off_b, stream); cudaFreeAsync(dense_a, stream); cudaFreeAsync(dense_b, stream);
    cudaFreeAsync(prefix, stream); if (dense_bases) cudaFreeAsync(dense_bases, stream);
    cudaFreeAsync(partial2, stream); cudaFreeAsync(items2, stream);
    cudaFreeAsync(partial, stream); cudaFreeAsync(red_a, stream); cudaFreeAsync(red_b, stream); cudaFreeAsync(cub_tmp, stream);
    return rc;
}

int msm_window_sums_device(uint32_t* d_window_sums, const MsmPlan& plan, const void* d_points, size_t stride,
                           const void* d_scalars, size_t npoints, cudaStream_t stream) {
    return msm_core(d_window_sums, plan, d_points, stride, nullptr, 0, d_scalars, npoints, stream);
}

MsmPlan msm_make_plan_precomputed(size_t npoints) {
    MsmPlan p;
    int lg = ceil_log2(npoints < 2 ? 2 : npoints);
    // one bucket set for all windows ⇒ the bucket reduction is paid once and wide windows are cheap: c = 22 at 2^24
    // (12 windows instead of 15).  Swept on a B200 with tools/tune_precomputed.py (profiles/tune_precomputed_r1.log).
    int c = lg >= 24 ? 22 : lg >= 22 ? lg - 2 : lg >= 20 ? lg - 3 : lg >= 8 ? lg - 2 : 6;
    if (const char* e = getenv("SNARKVM_B200_MSM_PRE_C")) { int v = atoi(e); if (v >= 2 && v <= 24) c = v; }
    p.c = c;
    p.nwin = 253 / c + 1;
    p.nbuckets = 1u << (c - 1);
    size_t total = npoints * (size_t)p.nwin;
    size_t cap = total / 300000 + 1;
    if (cap < 16) cap = 16;
    p.cap = (uint32_t)cap;
    int levels = lg >= 24 ? 3 : lg >= 22 ? 2 : 1;
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

int msm_precomputed_sum_device(uint32_t* d_sum, const MsmPlan& plan, const uint32_t* d_table, size_t table_n, const void* d_scalars,
                               size_t nscalars, cudaStream_t stream) {
    return msm_core(d_sum, plan, nullptr, 0, d_table, table_n, d_scalars, nscalars, stream);
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

int msm_generate_bases_device(void* d_points, size_t npoints, size_t stride, uint64_t seed, cudaStream_t stream) {
    if (stride < 104 || (stride & 7)) return (int)cudaErrorInvalidValue;
    if (npoints == 0) return 0;
    k_generate_bases<<<(unsigned)((npoints + 127) / 128), 128, 0, stream>>>((uint8_t*)d_points, npoints, stride, seed);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace b200
