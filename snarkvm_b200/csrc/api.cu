// C ABI (include/snarkvm_b200.h): the three drop-in symbols of the reference FFI
// (/root/reference/algorithms/cuda/src/lib.rs:42-69) plus the device-resident extended API.
// Error contract: never throw or abort across the boundary; return a
// cudaError_t code; leave outputs untouched on failure so the Rust caller can fall back.
#include "../../include/snarkvm_b200.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "host_ec.hpp"
#include "msm.cuh"
#include "ntt.cuh"
#include "poly.cuh"

using namespace b200;

namespace {

snarkvm_error_t make_error(int code) {
    snarkvm_error_t e;
    e.code = code;
    e.message = nullptr;
    if (code != 0) {
        const char* s = cudaGetErrorString((cudaError_t)code);
        if (s) { size_t n = strlen(s) + 1; e.message = (char*)malloc(n); if (e.message) memcpy(e.message, s, n); }
    }
    return e;
}

// One non-blocking stream per (host thread, device): the FFI is entered concurrently from many
// rayon workers (sonic_pc/mod.rs:186-245), so calls must not serialise on the default stream.
struct ThreadCtx {
    cudaStream_t stream[64] = {};
    bool have[64] = {};
    cudaStream_t copy_stream[64] = {};
    bool have_copy[64] = {};
    ~ThreadCtx() {}
};
thread_local ThreadCtx t_ctx;

int thread_stream(cudaStream_t* out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    dev &= 63;
    ensure_pool_configured();
    if (!t_ctx.have[dev]) {
        e = cudaStreamCreateWithFlags(&t_ctx.stream[dev], cudaStreamNonBlocking);
        if (e != cudaSuccess) return (int)e;
        t_ctx.have[dev] = true;
    }
    *out = t_ctx.stream[dev];
    return 0;
}

// second stream of the calling thread: uploads that overlap the kernels of the first
int thread_copy_stream(cudaStream_t* out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    dev &= 63;
    if (!t_ctx.have_copy[dev]) {
        e = cudaStreamCreateWithFlags(&t_ctx.copy_stream[dev], cudaStreamNonBlocking);
        if (e != cudaSuccess) return (int)e;
        t_ctx.have_copy[dev] = true;
    }
    *out = t_ctx.copy_stream[dev];
    return 0;
}

// Registered (resident) bases: the SRS powers of a proving key are constant, and the reference re-passes the same
// host slice on every commitment (kzg10/mod.rs:119,149).  A caller may register that slice once; snarkvm_msm then
// recognises the pointer and skips the 104 B/point upload.  Opt-in: the caller promises not to mutate the slice.
struct ResidentBases { void* d_ptr; size_t npoints; size_t stride; int device; void* tables; /* PrecomputedBases* or null */ };
std::mutex g_bases_mu;
std::map<const void*, ResidentBases> g_bases;

bool find_resident(const void* host, size_t npoints, size_t stride, ResidentBases* out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lock(g_bases_mu);
    auto it = g_bases.find(host);
    if (it == g_bases.end() || it->second.device != dev || it->second.stride != stride || it->second.npoints < npoints) return false;
    *out = it->second;
    return true;
}

void write_infinity(void* out144) {
    host::Xyzz inf = host::xyzz_inf();
    host::xyzz_to_normalised_projective(inf, (uint64_t*)out144);
}

int msm_device_impl(void* out144, const void* d_points, size_t npoints, const void* d_scalars, size_t stride,
                    cudaStream_t stream) {
    if (npoints == 0) { write_infinity(out144); return 0; }
    if (stride < 104 || (stride & 7)) return (int)cudaErrorInvalidValue;
    ensure_pool_configured();
    MsmPlan plan = msm_make_plan(npoints);
    uint32_t* d_sums = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&d_sums, (size_t)plan.nwin * 192, stream);
    if (e != cudaSuccess) return (int)e;
    int rc = msm_window_sums_device(d_sums, plan, d_points, stride, d_scalars, npoints, stream);
    std::vector<host::Xyzz> sums((size_t)plan.nwin);
    if (rc == 0) rc = (int)cudaMemcpyAsync(sums.data(), d_sums, (size_t)plan.nwin * 192, cudaMemcpyDeviceToHost, stream);
    cudaFreeAsync(d_sums, stream);
    if (rc == 0) rc = (int)cudaStreamSynchronize(stream);
    if (rc != 0) return rc;
    host::Xyzz total = host::horner_windows(sums.data(), plan.nwin, plan.c);
    host::xyzz_to_normalised_projective(total, (uint64_t*)out144);
    return 0;
}

// NN is native (bit reversal fused into the last pass); the other orders add explicit derange passes:
// R-input ⇒ permute before, R-output ⇒ permute after.
int ntt_ordered(void* d, uint32_t lg, int order, int dir, int type, void* scratch, cudaStream_t stream) {
    if (order < 0 || order > 3) return (int)cudaErrorInvalidValue;
    int rc = 0;
    if (order == SNARKVM_NTT_RN || order == SNARKVM_NTT_RR) rc = fr_bitrev_device(d, lg, stream);
    if (rc == 0) rc = ntt_device(d, lg, dir, type, scratch, stream);
    if (rc == 0 && (order == SNARKVM_NTT_NR || order == SNARKVM_NTT_RR)) rc = fr_bitrev_device(d, lg, stream);
    return rc;
}

int polymul_device_impl(void* d_out, size_t pcount, const void* const* d_polys, const size_t* plens, size_t ecount,
                        const void* const* d_evals, const size_t* elens, uint32_t lg, cudaStream_t stream) {
    ensure_pool_configured();
    const size_t n = (size_t)1 << lg, bytes = n * 32;
    if (pcount + ecount == 0) return 0;
    for (size_t i = 0; i < pcount; i++) if (plens[i] > n) return (int)cudaErrorInvalidValue;
    for (size_t i = 0; i < ecount; i++) if (elens[i] != n) return (int)cudaErrorInvalidValue;   // polynomial.cuh:95
    void *tmp = nullptr, *scratch = nullptr;
    cudaError_t e;
    if ((e = cudaMallocAsync(&tmp, bytes, stream)) != cudaSuccess) return (int)e;
    if ((e = cudaMallocAsync(&scratch, bytes, stream)) != cudaSuccess) { cudaFreeAsync(tmp, stream); return (int)e; }
    int rc = 0;
    bool have = false;
    for (size_t i = 0; i < pcount && rc == 0; i++) {
        void* dst = have ? tmp : d_out;
        rc = (int)cudaMemsetAsync(dst, 0, bytes, stream);
        if (rc == 0 && plens[i]) rc = (int)cudaMemcpyAsync(dst, d_polys[i], plens[i] * 32, cudaMemcpyDeviceToDevice, stream);
        if (rc == 0) rc = ntt_device(dst, lg, NTT_FORWARD, NTT_STANDARD, scratch, stream);
        if (rc == 0 && have) rc = fr_pointwise_mul_device(d_out, tmp, n, stream);
        have = true;
    }
    for (size_t i = 0; i < ecount && rc == 0; i++) {
        if (have) rc = fr_pointwise_mul_device(d_out, d_evals[i], n, stream);
        else rc = (int)cudaMemcpyAsync(d_out, d_evals[i], bytes, cudaMemcpyDeviceToDevice, stream);
        have = true;
    }
    if (rc == 0) rc = ntt_device(d_out, lg, NTT_INVERSE, NTT_STANDARD, scratch, stream);
    cudaFreeAsync(tmp, stream);
    cudaFreeAsync(scratch, stream);
    return rc;
}

}  // namespace

extern "C" {

const char* snarkvm_b200_version(void) { return "snarkvm_b200 0.1 (sm_100a)"; }
uint64_t snarkvm_b200_launch_count(void) { return launch_count(); }

// ----------------------------------------------------------------------------------------
// PART 1 — drop-in symbols
// ----------------------------------------------------------------------------------------
snarkvm_error_t snarkvm_ntt(void* inout, uint32_t lg, snarkvm_ntt_order_t order, snarkvm_ntt_direction_t dir,
                            snarkvm_ntt_type_t type) {
    if (lg > NTT_MAX_LG || !inout || (int)order < 0 || (int)order > 3) return make_error((int)cudaErrorInvalidValue);
    cudaStream_t stream;
    int rc = thread_stream(&stream);
    if (rc) return make_error(rc);
    const size_t bytes = ((size_t)1 << lg) * 32;
    void *d = nullptr, *scratch = nullptr;
    rc = (int)cudaMallocAsync(&d, bytes, stream);
    if (rc == 0) rc = (int)cudaMallocAsync(&scratch, bytes, stream);
    if (rc == 0) rc = (int)cudaMemcpyAsync(d, inout, bytes, cudaMemcpyHostToDevice, stream);
    if (rc == 0) rc = ntt_ordered(d, lg, (int)order, (int)dir, (int)type, scratch, stream);
    if (rc == 0) rc = (int)cudaStreamSynchronize(stream);          // only copy back on success (snarkvm.cu:178-183)
    if (rc == 0) rc = (int)cudaMemcpyAsync(inout, d, bytes, cudaMemcpyDeviceToHost, stream);
    if (d) cudaFreeAsync(d, stream);
    if (scratch) cudaFreeAsync(scratch, stream);
    int rs = (int)cudaStreamSynchronize(stream);
    return make_error(rc ? rc : rs);
}

snarkvm_error_t snarkvm_polymul(void* out, size_t pcount, const void* polynomials, const void* plens_v, size_t ecount,
                                const void* evaluations, const void* elens_v, uint32_t lg) {
    if (pcount + ecount == 0) return make_error(0);                 // snarkvm.cu:195-197
    if (lg > NTT_MAX_LG || !out) return make_error((int)cudaErrorInvalidValue);
    const void* const* polys = (const void* const*)polynomials;
    const void* const* evals = (const void* const*)evaluations;
    const size_t* plens = (const size_t*)plens_v;
    const size_t* elens = (const size_t*)elens_v;
    const size_t n = (size_t)1 << lg, bytes = n * 32;
    for (size_t i = 0; i < pcount; i++) if (plens[i] > n) return make_error((int)cudaErrorInvalidValue);
    for (size_t i = 0; i < ecount; i++) if (elens[i] != n) return make_error((int)cudaErrorInvalidValue);
    cudaStream_t stream;
    int rc = thread_stream(&stream);
    if (rc) return make_error(rc);
    std::vector<void*> dbuf(pcount + ecount, nullptr);
    std::vector<const void*> dp(pcount), de(ecount);
    void* d_out = nullptr;
    rc = (int)cudaMallocAsync(&d_out, bytes, stream);
    for (size_t i = 0; i < pcount && rc == 0; i++) {
        size_t b = plens[i] ? plens[i] * 32 : 32;
        rc = (int)cudaMallocAsync(&dbuf[i], b, stream);
        if (rc == 0 && plens[i]) rc = (int)cudaMemcpyAsync(dbuf[i], polys[i], plens[i] * 32, cudaMemcpyHostToDevice, stream);
        dp[i] = dbuf[i];
    }
    for (size_t i = 0; i < ecount && rc == 0; i++) {
        rc = (int)cudaMallocAsync(&dbuf[pcount + i], bytes, stream);
        if (rc == 0) rc = (int)cudaMemcpyAsync(dbuf[pcount + i], evals[i], bytes, cudaMemcpyHostToDevice, stream);
        de[i] = dbuf[pcount + i];
    }
    if (rc == 0) rc = polymul_device_impl(d_out, pcount, dp.data(), plens, ecount, de.data(), elens, lg, stream);
    if (rc == 0) rc = (int)cudaStreamSynchronize(stream);
    if (rc == 0) rc = (int)cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, stream);
    for (void* p : dbuf) if (p) cudaFreeAsync(p, stream);
    if (d_out) cudaFreeAsync(d_out, stream);
    int rs = (int)cudaStreamSynchronize(stream);
    return make_error(rc ? rc : rs);
}

// Large host-buffer MSMs are cut into point ranges: range k+1 crosses PCIe on the thread's copy stream while range k is in
// the Pippenger kernels (an MSM is a sum over points: every range is a complete MSM with its own plan and the results add).
// 2^24 points: 2.28 GB of upload, ≈ 40 ms, of which only the first range's share stays exposed — so the first range is small
// (1/8 of the points), then 3/8, then 1/2.  SNARKVM_B200_MSM_CHUNKS = "1", "2" (equal parts) or weights like "1:3:4".
static std::vector<size_t> msm_ranges(size_t npoints) {
    std::vector<size_t> w;
    if (npoints >= ((size_t)1 << 23)) w = {1, 3, 4};
    if (const char* e = getenv("SNARKVM_B200_MSM_CHUNKS")) {
        std::vector<size_t> v;
        for (const char* p = e; *p;) {
            char* end = nullptr;
            long x = strtol(p, &end, 10);
            if (end == p || x < 1 || x > 1024) { v.clear(); break; }
            v.push_back((size_t)x);
            p = *end == ':' ? end + 1 : end;
            if (*end && *end != ':') { v.clear(); break; }
        }
        if (v.size() == 1) v.assign(v[0] <= 64 ? v[0] : 64, 1);            // "k" = k equal parts
        if (!v.empty()) w = v;
    }
    size_t total = 0;
    for (size_t x : w) total += x;
    std::vector<size_t> bounds{0};                                        // range k = [bounds[k], bounds[k+1])
    size_t acc = 0;
    for (size_t x : w) {
        acc += x;
        size_t b = (size_t)((unsigned __int128)npoints * acc / total);
        if (b > bounds.back()) bounds.push_back(b);                       // empty ranges are dropped
    }
    if (bounds.back() != npoints) bounds.push_back(npoints);
    return bounds;
}

snarkvm_error_t snarkvm_msm(void* out, const void* points, size_t npoints, const void* scalars, size_t ffi_affine_sz) {
    if (!out) return make_error((int)cudaErrorInvalidValue);
    if (npoints == 0) { write_infinity(out); return make_error(0); }
    if (!points || !scalars || ffi_affine_sz < 104 || (ffi_affine_sz & 7)) return make_error((int)cudaErrorInvalidValue);
    cudaStream_t stream;
    int rc = thread_stream(&stream);
    if (rc) return make_error(rc);
    void *d_points = nullptr, *d_scalars = nullptr;
    ResidentBases rb;
    const bool resident = find_resident(points, npoints, ffi_affine_sz, &rb);
    std::vector<size_t> bounds{0, npoints};
    if (!resident) bounds = msm_ranges(npoints);
    const int chunks = (int)bounds.size() - 1;
    if (!resident) rc = (int)cudaMallocAsync(&d_points, npoints * ffi_affine_sz, stream);
    if (rc == 0) rc = (int)cudaMallocAsync(&d_scalars, npoints * 32, stream);
    uint64_t result[18];
    if (chunks == 1) {
        if (rc == 0) rc = (int)cudaMemcpyAsync(d_scalars, scalars, npoints * 32, cudaMemcpyHostToDevice, stream);
        if (rc == 0 && !resident) rc = (int)cudaMemcpyAsync(d_points, points, npoints * ffi_affine_sz, cudaMemcpyHostToDevice, stream);
        if (rc == 0 && resident && rb.tables) rc = snarkvm_b200_msm_precomputed_device(result, rb.tables, d_scalars, npoints, stream);
        else if (rc == 0) rc = msm_device_impl(result, resident ? rb.d_ptr : d_points, npoints, d_scalars, ffi_affine_sz, stream);
    } else {
        cudaStream_t copy = nullptr;
        if (rc == 0) rc = thread_copy_stream(&copy);
        std::vector<MsmPlan> plans;
        std::vector<size_t> sum_off{0};                                   // in XYZZ points
        for (int k = 0; k < chunks; k++) {
            plans.push_back(msm_make_plan(bounds[k + 1] - bounds[k]));
            sum_off.push_back(sum_off.back() + (size_t)plans.back().nwin);
        }
        std::vector<cudaEvent_t> ev((size_t)chunks + 1, nullptr);
        for (auto& e : ev) if (rc == 0) rc = (int)cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        uint32_t* d_sums = nullptr;
        if (rc == 0) rc = (int)cudaMallocAsync((void**)&d_sums, sum_off.back() * 192, stream);
        // the pool allocations above are ordered on `stream`; the copy stream may touch them only after this point
        if (rc == 0) rc = (int)cudaEventRecord(ev[chunks], stream);
        if (rc == 0) rc = (int)cudaStreamWaitEvent(copy, ev[chunks], 0);
        for (int k = 0; k < chunks && rc == 0; k++) {
            const size_t i0 = bounds[k], cn = bounds[k + 1] - i0;
            rc = (int)cudaMemcpyAsync((uint8_t*)d_scalars + i0 * 32, (const uint8_t*)scalars + i0 * 32, cn * 32, cudaMemcpyHostToDevice, copy);
            if (rc == 0) rc = (int)cudaMemcpyAsync((uint8_t*)d_points + i0 * ffi_affine_sz, (const uint8_t*)points + i0 * ffi_affine_sz,
                                                   cn * ffi_affine_sz, cudaMemcpyHostToDevice, copy);
            if (rc == 0) rc = (int)cudaEventRecord(ev[k], copy);
        }
        for (int k = 0; k < chunks && rc == 0; k++) {
            const size_t i0 = bounds[k], cn = bounds[k + 1] - i0;
            rc = (int)cudaStreamWaitEvent(stream, ev[k], 0);
            if (rc == 0) rc = msm_window_sums_device(d_sums + sum_off[k] * 48, plans[k], (const uint8_t*)d_points + i0 * ffi_affine_sz,
                                                     ffi_affine_sz, (const uint8_t*)d_scalars + i0 * 32, cn, stream);
        }
        std::vector<host::Xyzz> sums(sum_off.back());
        if (rc == 0) rc = (int)cudaMemcpyAsync(sums.data(), d_sums, sums.size() * 192, cudaMemcpyDeviceToHost, stream);
        if (d_sums) cudaFreeAsync(d_sums, stream);
        // on any failure the copy stream may still be writing: drain it before the buffers go back to the pool
        if (copy) { int rs = (int)cudaStreamSynchronize(copy); if (rc == 0) rc = rs; }
        if (rc == 0) rc = (int)cudaStreamSynchronize(stream);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (rc == 0) {
            host::Xyzz total = host::xyzz_inf();
            for (int k = 0; k < chunks; k++) host::xyzz_add(total, host::horner_windows(sums.data() + sum_off[k], plans[k].nwin, plans[k].c));
            host::xyzz_to_normalised_projective(total, result);
        }
    }
    if (d_points) cudaFreeAsync(d_points, stream);
    if (d_scalars) cudaFreeAsync(d_scalars, stream);
    int rs = (int)cudaStreamSynchronize(stream);
    if (rc == 0 && rs == 0) memcpy(out, result, 144);
    return make_error(rc ? rc : rs);
}

// ----------------------------------------------------------------------------------------
// PART 2 — extended device-resident API
// ----------------------------------------------------------------------------------------
int snarkvm_b200_ntt_device(void* d_inout, uint32_t lg, int order, int dir, int type, void* d_scratch, void* stream) {
    if (lg > NTT_MAX_LG) return (int)cudaErrorInvalidValue;
    return ntt_ordered(d_inout, lg, order, dir, type, d_scratch, (cudaStream_t)stream);
}

int snarkvm_b200_polymul_device(void* d_out, size_t pcount, const void* const* d_polys, const size_t* plens, size_t ecount,
                                const void* const* d_evals, const size_t* elens, uint32_t lg, void* stream) {
    if (lg > NTT_MAX_LG) return (int)cudaErrorInvalidValue;
    return polymul_device_impl(d_out, pcount, d_polys, plens, ecount, d_evals, elens, lg, (cudaStream_t)stream);
}

int snarkvm_b200_msm_plan(size_t npoints, int* c, int* nwin, uint32_t* cap) {
    MsmPlan p = msm_make_plan(npoints);
    if (c) *c = p.c;
    if (nwin) *nwin = p.nwin;
    if (cap) *cap = p.cap;
    return 0;
}

int snarkvm_b200_msm_device(void* out144, const void* d_points, size_t npoints, const void* d_scalars, size_t stride,
                            void* stream) {
    if (!out144) return (int)cudaErrorInvalidValue;
    return msm_device_impl(out144, d_points, npoints, d_scalars, stride, (cudaStream_t)stream);
}

int snarkvm_b200_msm_window_sums_device(void* d_window_sums, const void* d_points, size_t npoints, const void* d_scalars,
                                        size_t stride, void* stream) {
    if (stride < 104 || (stride & 7) || npoints == 0) return (int)cudaErrorInvalidValue;
    MsmPlan plan = msm_make_plan(npoints);
    return msm_window_sums_device((uint32_t*)d_window_sums, plan, d_points, stride, d_scalars, npoints, (cudaStream_t)stream);
}

int snarkvm_b200_xyzz_sum_ranks_device(void* d_out, const void* d_in, int nranks, int count, void* stream) {
    if (nranks < 1 || count < 1) return (int)cudaErrorInvalidValue;
    return xyzz_sum_ranks_device((uint32_t*)d_out, (const uint32_t*)d_in, nranks, count, (cudaStream_t)stream);
}

int snarkvm_b200_msm_finish(void* out144, const void* h_window_sums, int nwin, int c) {
    if (!out144 || !h_window_sums || nwin < 1 || c < 0) return (int)cudaErrorInvalidValue;
    std::vector<host::Xyzz> sums((size_t)nwin);
    memcpy(sums.data(), h_window_sums, (size_t)nwin * 192);
    host::Xyzz total = host::horner_windows(sums.data(), nwin, c);
    host::xyzz_to_normalised_projective(total, (uint64_t*)out144);
    return 0;
}

int snarkvm_b200_kzg_commit_device(void* out144, const void* d_powers, size_t stride, const void* d_coeffs_mont,
                                   size_t ncoeffs, void* stream_v) {
    if (!out144) return (int)cudaErrorInvalidValue;
    if (ncoeffs == 0) { write_infinity(out144); return 0; }
    cudaStream_t stream = (cudaStream_t)stream_v;
    void* d_plain = nullptr;
    int rc = (int)cudaMallocAsync(&d_plain, ncoeffs * 32, stream);
    if (rc == 0) rc = fr_from_mont_device(d_plain, d_coeffs_mont, ncoeffs, stream);
    if (rc == 0) rc = msm_device_impl(out144, d_powers, ncoeffs, d_plain, stride, stream);
    if (d_plain) cudaFreeAsync(d_plain, stream);
    return rc;
}

// Precomputed tables for a fixed base set (an SRS): handle = {table, n, plan}.  One-time cost: (nwin−1)·c doublings and
// nwin−1 inversions per point; memory npoints·nwin·128 B.
struct PrecomputedBases { uint32_t* table; size_t n; MsmPlan plan; int device; };

int snarkvm_b200_msm_precompute_device(void** handle_out, const void* d_points, size_t npoints, size_t stride, void* stream_v) {
    if (!handle_out || !d_points || npoints == 0 || stride < 104 || (stride & 7)) return (int)cudaErrorInvalidValue;
    cudaStream_t stream = (cudaStream_t)stream_v;
    MsmPlan plan = msm_make_plan_precomputed(npoints);
    if (npoints * (size_t)plan.nwin >= (1ull << 31)) return (int)cudaErrorInvalidValue;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    uint32_t* table = nullptr;
    if ((e = cudaMalloc(&table, npoints * (size_t)plan.nwin * 128)) != cudaSuccess) return (int)e;
    int rc = msm_precompute_tables_device(table, plan, d_points, stride, npoints, stream);
    if (rc == 0) rc = (int)cudaStreamSynchronize(stream);
    if (rc != 0) { cudaFree(table); return rc; }
    *handle_out = new PrecomputedBases{table, npoints, plan, dev};
    return 0;
}

int snarkvm_b200_msm_precomputed_free(void* handle) {
    if (!handle) return (int)cudaErrorInvalidValue;
    PrecomputedBases* h = (PrecomputedBases*)handle;
    cudaError_t e = cudaFree(h->table);
    delete h;
    return (int)e;
}

int snarkvm_b200_msm_precomputed_info(const void* handle, size_t* npoints, int* c, int* nwin, size_t* table_bytes) {
    if (!handle) return (int)cudaErrorInvalidValue;
    const PrecomputedBases* h = (const PrecomputedBases*)handle;
    if (npoints) *npoints = h->n;
    if (c) *c = h->plan.c;
    if (nwin) *nwin = h->plan.nwin;
    if (table_bytes) *table_bytes = h->n * (size_t)h->plan.nwin * 128;
    return 0;
}

int snarkvm_b200_msm_precomputed_device(void* out144, const void* handle, const void* d_scalars, size_t nscalars, void* stream_v) {
    if (!out144 || !handle) return (int)cudaErrorInvalidValue;
    const PrecomputedBases* h = (const PrecomputedBases*)handle;
    if (nscalars > h->n) return (int)cudaErrorInvalidValue;
    if (nscalars == 0) { write_infinity(out144); return 0; }
    if (!d_scalars) return (int)cudaErrorInvalidValue;
    cudaStream_t stream = (cudaStream_t)stream_v;
    uint32_t* d_sum = nullptr;
    cudaError_t e = cudaMallocAsync(&d_sum, 192, stream);
    if (e != cudaSuccess) return (int)e;
    int rc = msm_precomputed_sum_device(d_sum, h->plan, h->table, h->n, d_scalars, nscalars, stream);
    host::Xyzz sum;
    if (rc == 0) rc = (int)cudaMemcpyAsync(&sum, d_sum, 192, cudaMemcpyDeviceToHost, stream);
    cudaFreeAsync(d_sum, stream);
    if (rc == 0) rc = (int)cudaStreamSynchronize(stream);
    if (rc != 0) return rc;
    host::xyzz_to_normalised_projective(sum, (uint64_t*)out144);
    return 0;
}

int snarkvm_b200_kzg_commit_precomputed_device(void* out144, const void* handle, const void* d_coeffs_mont, size_t ncoeffs, void* stream_v) {
    if (!out144 || !handle) return (int)cudaErrorInvalidValue;
    if (ncoeffs == 0) { write_infinity(out144); return 0; }
    cudaStream_t stream = (cudaStream_t)stream_v;
    void* d_plain = nullptr;
    int rc = (int)cudaMallocAsync(&d_plain, ncoeffs * 32, stream);
    if (rc == 0) rc = fr_from_mont_device(d_plain, d_coeffs_mont, ncoeffs, stream);
    if (rc == 0) rc = snarkvm_b200_msm_precomputed_device(out144, handle, d_plain, ncoeffs, stream_v);
    if (d_plain) cudaFreeAsync(d_plain, stream);
    return rc;
}

// KZG10::commit with a hiding bound (polycommit/kzg10/mod.rs:98-156): commitment to the plaintext polynomial against
// powers_of_beta_g plus the commitment to the blinding polynomial against powers_of_beta_times_gamma_g.  The caller samples the
// blinding polynomial (KZGRandomness::rand, :129-140) and passes its Montgomery coefficients; nblinding = 0 is the non-hiding
// commit.  Zero coefficients contribute nothing, so skip_leading_zeros_and_convert_to_bigints (:455-467) needs no special path.
int snarkvm_b200_kzg_commit_hiding_device(void* out144, const void* d_powers, size_t stride, const void* d_coeffs_mont, size_t ncoeffs,
                                          const void* d_gamma_powers, const void* d_blinding_mont, size_t nblinding, void* stream) {
    if (!out144) return (int)cudaErrorInvalidValue;
    uint64_t a[18], b[18];
    int rc = snarkvm_b200_kzg_commit_device(a, d_powers, stride, d_coeffs_mont, ncoeffs, stream);
    if (rc == 0) rc = snarkvm_b200_kzg_commit_device(b, d_gamma_powers, stride, d_blinding_mont, nblinding, stream);
    if (rc != 0) return rc;
    host::Xyzz sum = host::xyzz_from_projective(a);
    host::xyzz_add(sum, host::xyzz_from_projective(b));
    host::xyzz_to_normalised_projective(sum, (uint64_t*)out144);
    return 0;
}

// All commitments of one prover round share powers_of_beta_g (sonic_pc/mod.rs:177-257): count polynomials, one call, the
// bases stay where they are.  out144s: count × 144 B of HOST memory.
int snarkvm_b200_kzg_commit_batch_device(void* out144s, const void* d_powers, size_t stride, const void* const* d_coeffs_mont,
                                         const size_t* ncoeffs, size_t count, void* stream) {
    if (count == 0) return 0;
    if (!out144s || !d_coeffs_mont || !ncoeffs) return (int)cudaErrorInvalidValue;
    for (size_t i = 0; i < count; i++) {
        int rc = snarkvm_b200_kzg_commit_device((uint8_t*)out144s + i * 144, d_powers, stride, d_coeffs_mont[i], ncoeffs[i], stream);
        if (rc != 0) return rc;
    }
    return 0;
}

int snarkvm_b200_g1_ntt_device(void* d_out, size_t out_stride, const void* d_in, size_t in_stride, uint32_t lg, int direction, void* stream) {
    return g1_ntt_device(d_out, out_stride, d_in, in_stride, lg, direction, (cudaStream_t)stream);
}
int snarkvm_b200_fr_batch_inversion_and_mul_device(void* d_v, size_t n, const void* coeff_mont_host, void* stream) {
    return fr_batch_inversion_and_mul_device(d_v, n, coeff_mont_host, (cudaStream_t)stream);
}
int snarkvm_b200_poly_divide_by_vanishing_device(void* d_q, void* d_r, const void* d_p, size_t m, size_t n, void* stream) {
    return poly_divide_by_vanishing_device(d_q, d_r, d_p, m, n, (cudaStream_t)stream);
}
int snarkvm_b200_poly_divide_by_linear_device(void* d_q, const void* d_p, size_t m, const void* point_mont_host, void* stream) {
    return poly_divide_by_linear_device(d_q, d_p, m, point_mont_host, (cudaStream_t)stream);
}
int snarkvm_b200_sparse_matvec_device(void* d_out, const void* d_row_ptr, const void* d_cols, const void* d_vals, size_t nrows, const void* d_x,
                                      size_t nvars, void* stream) {
    return sparse_matvec_device(d_out, d_row_ptr, d_cols, d_vals, nrows, d_x, nvars, (cudaStream_t)stream);
}
int snarkvm_b200_fr_vec_op_device(void* d_out, const void* d_a, const void* d_b, size_t n, int op, void* stream) {
    return fr_vec_op_device(d_out, d_a, d_b, n, op, (cudaStream_t)stream);
}
int snarkvm_b200_fr_vec_scalar_op_device(void* d_out, const void* d_a, const void* scalar_mont_host, size_t n, int op, void* stream) {
    return fr_vec_scalar_op_device(d_out, d_a, scalar_mont_host, n, op, (cudaStream_t)stream);
}
int snarkvm_b200_domain_elements_device(void* d_out, uint32_t lg, void* stream) {
    return domain_elements_device(d_out, lg, (cudaStream_t)stream);
}
int snarkvm_b200_poly_evaluate_device(void* out_mont_host, const void* d_coeffs, size_t m, const void* point_mont_host, void* stream) {
    return poly_evaluate_device(out_mont_host, d_coeffs, m, point_mont_host, (cudaStream_t)stream);
}

int snarkvm_b200_fr_from_mont_device(void* d_out, const void* d_in, size_t n, void* stream) {
    return fr_from_mont_device(d_out, d_in, n, (cudaStream_t)stream);
}
int snarkvm_b200_fr_to_mont_device(void* d_out, const void* d_in, size_t n, void* stream) {
    return fr_to_mont_device(d_out, d_in, n, (cudaStream_t)stream);
}

int snarkvm_b200_srs_decode_device(void* d_out, size_t stride, const void* d_in96, size_t npoints, uint32_t* d_invalid, void* stream) {
    if (!d_out || !d_in96 || !d_invalid) return (int)cudaErrorInvalidValue;
    return srs_decode_device(d_out, stride, d_in96, npoints, d_invalid, (cudaStream_t)stream);
}

int snarkvm_b200_register_bases(const void* host_points, size_t npoints, size_t stride) {
    if (!host_points || npoints == 0 || stride < 104 || (stride & 7)) return (int)cudaErrorInvalidValue;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    void* d = nullptr;
    if ((e = cudaMalloc(&d, npoints * stride)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpy(d, host_points, npoints * stride, cudaMemcpyHostToDevice)) != cudaSuccess) { cudaFree(d); return (int)e; }
    std::lock_guard<std::mutex> lock(g_bases_mu);
    auto it = g_bases.find(host_points);
    if (it != g_bases.end()) { cudaFree(it->second.d_ptr); if (it->second.tables) snarkvm_b200_msm_precomputed_free(it->second.tables); g_bases.erase(it); }
    g_bases[host_points] = ResidentBases{d, npoints, stride, dev, nullptr};
    return 0;
}
int snarkvm_b200_unregister_bases(const void* host_points) {
    std::lock_guard<std::mutex> lock(g_bases_mu);
    auto it = g_bases.find(host_points);
    if (it == g_bases.end()) return (int)cudaErrorInvalidValue;
    cudaFree(it->second.d_ptr);
    if (it->second.tables) snarkvm_b200_msm_precomputed_free(it->second.tables);
    g_bases.erase(it);
    return 0;
}
// register + build the fixed-base tables (snarkvm_b200_msm_precompute_device) from the uploaded copy: later snarkvm_msm calls
// on this slice run over the tables (one bucket set, wider windows).  Costs npoints·nwin·128 B of HBM and seconds of set-up.
int snarkvm_b200_register_bases_precomputed(const void* host_points, size_t npoints, size_t stride) {
    int rc = snarkvm_b200_register_bases(host_points, npoints, stride);
    if (rc != 0) return rc;
    void* d = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_bases_mu);
        d = g_bases[host_points].d_ptr;
    }
    cudaStream_t stream;
    if ((rc = thread_stream(&stream)) != 0) return rc;
    void* tables = nullptr;
    rc = snarkvm_b200_msm_precompute_device(&tables, d, npoints, stride, stream);
    if (rc != 0) { snarkvm_b200_unregister_bases(host_points); return rc; }
    std::lock_guard<std::mutex> lock(g_bases_mu);
    auto it = g_bases.find(host_points);
    if (it == g_bases.end() || it->second.d_ptr != d) { snarkvm_b200_msm_precomputed_free(tables); return (int)cudaErrorInvalidValue; }   // raced with unregister
    it->second.tables = tables;
    return 0;
}

int snarkvm_b200_profile_enable(int on) { prof_enable(on != 0); return 0; }
int snarkvm_b200_profile_collect(int kind, double* total_ms, uint64_t* count) { return prof_collect(kind, total_ms, count); }

int snarkvm_b200_generate_bases_device(void* d_points, size_t npoints, size_t stride, uint64_t seed, void* stream) {
    return msm_generate_bases_device(d_points, npoints, stride, seed, (cudaStream_t)stream);
}

}  // extern "C"
