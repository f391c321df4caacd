// snarkvm_b200.hpp — C++ host-side mirror of the reference operator surface for the hot path, over the C ABI of
// snarkvm_b200.h.  The reference is Rust; no Rust toolchain exists in this image, so the compiled-language host side
// is C++ with the reference's names, argument meaning and error behaviour:
//
//   snarkvm_b200::cuda::{NTT, polymul, msm}      ↔ snarkvm_algorithms_cuda::{NTT, polymul, msm}
//                                                   (/root/reference/algorithms/cuda/src/lib.rs:71-168)
//   snarkvm_b200::VariableBase::msm               ↔ algorithms/src/msm/variable_base/mod.rs:27-49
//   snarkvm_b200::EvaluationDomain                ↔ algorithms/src/fft/domain.rs:82-221
//   snarkvm_b200::PolyMultiplier                  ↔ algorithms/src/fft/polynomial/multiplier.rs:28-134
//
// Error behaviour: the Rust shims return Result<_, cuda::Error> and `panic!` on caller bugs (lib.rs:84-86,150-152).
// Here Result<T> carries {code, message} (code 0 = Ok) and caller bugs throw std::invalid_argument.  There is no CPU
// fallback in this library: where the reference falls back to its CPU path on Err (variable_base/mod.rs:39-43,
// fft/domain.rs:385-387) the mirror returns the error to the caller.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "snarkvm_b200.h"

namespace snarkvm_b200 {

// In-memory images of the reference types (little-endian u64 limbs).
struct Fr { uint64_t l[4]; };                      // Fp256<FrParameters>, Montgomery form (fields/src/fp_256.rs:52)
struct BigInteger256 { uint64_t l[4]; };           // canonical scalar (utilities/src/biginteger/bigint_256.rs:36)
struct Fq { uint64_t l[6]; };                      // Fp384<FqParameters>, Montgomery form (fields/src/fp_384.rs:52)
struct G1Affine { Fq x, y; bool infinity; uint8_t pad[7]; };   // affine.rs:41-46 — 104 bytes
struct G1Projective { Fq x, y, z; };               // projective.rs:36-41 — 144 bytes, infinity ⇔ z == 0
struct Fq2 { Fq c0, c1; };                         // Fp2<Fq2Parameters> (curves/src/bls12_377/fq2.rs:29)
struct G2Affine { Fq2 x, y; bool infinity; uint8_t pad[7]; };  // the same Affine template over Fq2 — 200 bytes
struct G2Projective { Fq2 x, y, z; };              // 288 bytes
static_assert(sizeof(G1Affine) == 104 && sizeof(G1Projective) == 144 && sizeof(Fr) == 32, "reference layouts");

struct CudaError {                                  // cuda::Error (lib.rs:19)
    int code = 0;
    std::string message;
};
template <class T>
struct Result {
    CudaError err;
    T value{};
    bool is_ok() const { return err.code == 0; }
    bool is_err() const { return err.code != 0; }
};
struct Unit {};

namespace detail {
inline CudaError take(snarkvm_error_t e) {
    CudaError r;
    r.code = e.code;
    if (e.message) { r.message = e.message; std::free(e.message); }
    return r;
}
inline bool is_pow2(size_t v) { return v && !(v & (v - 1)); }
inline uint32_t log2_exact(size_t v) { uint32_t l = 0; while ((size_t(1) << l) < v) l++; return l; }
}  // namespace detail

namespace cuda {
enum class NTTInputOutputOrder : int { NN = 0, NR = 1, RN = 2, RR = 3 };   // lib.rs:22-28
enum class NTTDirection : int { Forward = 0, Inverse = 1 };                // lib.rs:30-34
enum class NTTType : int { Standard = 0, Coset = 1 };                      // lib.rs:36-40

// lib.rs:77-97
inline Result<Unit> NTT(size_t domain_size, Fr* inout, NTTInputOutputOrder order, NTTDirection direction, NTTType type) {
    if (!detail::is_pow2(domain_size)) throw std::invalid_argument("domain_size is not power of 2");
    Result<Unit> r;
    r.err = detail::take(snarkvm_ntt(inout, detail::log2_exact(domain_size), (snarkvm_ntt_order_t)order,
                                     (snarkvm_ntt_direction_t)direction, (snarkvm_ntt_type_t)type));
    return r;
}
// lib.rs:100-145
inline Result<std::vector<Fr>> polymul(size_t domain, const std::vector<std::vector<Fr>>& polynomials,
                                       const std::vector<std::vector<Fr>>& evaluations, const Fr& zero) {
    if (!detail::is_pow2(domain)) throw std::invalid_argument("domain_size is not power of 2");
    std::vector<const void*> pp, ep;
    std::vector<size_t> pl, el;
    for (auto& p : polynomials) { pp.push_back(p.data()); pl.push_back(p.size()); }
    for (auto& e : evaluations) { ep.push_back(e.data()); el.push_back(e.size()); }
    Result<std::vector<Fr>> r;
    r.value.assign(domain, zero);
    r.err = detail::take(snarkvm_polymul(r.value.data(), pp.size(), pp.data(), pl.data(), ep.size(), ep.data(), el.data(),
                                         detail::log2_exact(domain)));
    return r;
}
// lib.rs:148-168
inline Result<G1Projective> msm(const G1Affine* points, size_t npoints_available, const BigInteger256* scalars, size_t nscalars) {
    if (nscalars > npoints_available) throw std::invalid_argument("length mismatch: fewer points than scalars");
    Result<G1Projective> r;
    r.err = detail::take(snarkvm_msm(&r.value, points, nscalars, scalars, sizeof(G1Affine)));
    return r;
}
// the same call for Affine<G2> (the reference routes every curve but BLS12-377 G1 to standard::msm, mod.rs:44-47)
inline Result<G2Projective> msm_g2(const G2Affine* points, size_t npoints_available, const BigInteger256* scalars, size_t nscalars) {
    if (nscalars > npoints_available) throw std::invalid_argument("length mismatch: fewer points than scalars");
    Result<G2Projective> r;
    r.err = detail::take(snarkvm_b200_msm_g2(&r.value, points, nscalars, scalars, sizeof(G2Affine)));
    return r;
}
}  // namespace cuda

// algorithms/src/msm/variable_base/mod.rs:27-49.  The reference uses the GPU only for len > 1024 and otherwise (or on Err)
// its CPU path; this backend has no CPU path, every size runs on the device and errors are returned.
struct VariableBase {
    static Result<G1Projective> msm(const std::vector<G1Affine>& bases, const std::vector<BigInteger256>& scalars) {
        return cuda::msm(bases.data(), bases.size(), scalars.data(), scalars.size());
    }
    // the generic arm of the TypeId dispatch (mod.rs:44-47)
    static Result<G2Projective> msm(const std::vector<G2Affine>& bases, const std::vector<BigInteger256>& scalars) {
        return cuda::msm_g2(bases.data(), bases.size(), scalars.data(), scalars.size());
    }
};

// algorithms/src/fft/domain.rs:82-221
class EvaluationDomain {
  public:
    static constexpr uint32_t TWO_ADICITY = 47;    // curves/src/bls12_377/fr.rs:109
    // domain.rs:118-147
    static std::optional<EvaluationDomain> new_(size_t num_coeffs) {
        size_t size = 1;
        while (size < num_coeffs) { size <<= 1; if (!size) return std::nullopt; }
        uint32_t lg = detail::log2_exact(size);
        if (lg > TWO_ADICITY) return std::nullopt;
        return EvaluationDomain(size, lg);
    }
    static std::optional<size_t> compute_size_of_domain(size_t num_coeffs) {
        auto d = new_(num_coeffs);
        if (!d) return std::nullopt;
        return d->size();
    }
    size_t size() const { return size_; }
    uint32_t log_size_of_group() const { return lg_; }

    // `coeffs.resize(self.size(), T::zero())` then the in-order transform (domain.rs:169-221)
    Result<Unit> fft_in_place(std::vector<Fr>& coeffs) const { return run(coeffs, cuda::NTTDirection::Forward, cuda::NTTType::Standard); }
    Result<Unit> ifft_in_place(std::vector<Fr>& evals) const { return run(evals, cuda::NTTDirection::Inverse, cuda::NTTType::Standard); }
    Result<Unit> coset_fft_in_place(std::vector<Fr>& coeffs) const { return run(coeffs, cuda::NTTDirection::Forward, cuda::NTTType::Coset); }
    Result<Unit> coset_ifft_in_place(std::vector<Fr>& evals) const { return run(evals, cuda::NTTDirection::Inverse, cuda::NTTType::Coset); }

  private:
    EvaluationDomain(size_t size, uint32_t lg) : size_(size), lg_(lg) {}
    Result<Unit> run(std::vector<Fr>& x, cuda::NTTDirection d, cuda::NTTType t) const {
        if (x.size() > size_) throw std::invalid_argument("more coefficients than the domain size");
        x.resize(size_, Fr{{0, 0, 0, 0}});
        return cuda::NTT(size_, x.data(), cuda::NTTInputOutputOrder::NN, d, t);
    }
    size_t size_;
    uint32_t lg_;
};

// algorithms/src/fft/polynomial/multiplier.rs:28-134
class PolyMultiplier {
  public:
    void add_polynomial(std::vector<Fr> p, const std::string& = "") { polynomials_.push_back(std::move(p)); }
    void add_evaluation(std::vector<Fr> e, const std::string& = "") { evaluations_.push_back(std::move(e)); }
    // nullopt when there is nothing to multiply or an evaluation's length differs from the domain (multiplier.rs:70-78)
    std::optional<Result<std::vector<Fr>>> multiply() const {
        if (polynomials_.empty() && evaluations_.empty()) return std::nullopt;
        size_t degree = 0;
        for (auto& p : polynomials_) degree += p.size();
        auto domain = EvaluationDomain::new_(polynomials_.empty() ? evaluations_[0].size() : degree);
        if (!domain) return std::nullopt;
        for (auto& e : evaluations_) if (e.size() != domain->size()) return std::nullopt;
        return cuda::polymul(domain->size(), polynomials_, evaluations_, Fr{{0, 0, 0, 0}});
    }

  private:
    std::vector<std::vector<Fr>> polynomials_, evaluations_;
};

}  // namespace snarkvm_b200
