// Test driver for include/snarkvm_b200.hpp: reads binary inputs written by tests/test_cpp_mirror.py, calls the C++
// mirror (VariableBase::msm, EvaluationDomain, PolyMultiplier) and writes the results back for comparison with the oracle.
//   mirror_driver <dir>         — dir holds bases.bin scalars.bin fr.bin p1.bin p2.bin
// Exit code: 0 = all calls Ok, 3 = a call returned a CUDA error (printed), 2 = usage / IO error.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "snarkvm_b200.hpp"

using namespace snarkvm_b200;

template <class T> static std::vector<T> read_all(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::cerr << "cannot open " << path << "\n"; std::exit(2); }
    size_t bytes = (size_t)f.tellg();
    std::vector<T> v(bytes / sizeof(T));
    f.seekg(0);
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
    return v;
}
template <class T> static void write_all(const std::string& path, const T* data, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(data), (std::streamsize)(n * sizeof(T)));
}
template <class R> static void check(const R& r, const char* what) {
    if (r.is_err()) { std::cerr << what << ": cuda error " << r.err.code << " " << r.err.message << "\n"; std::exit(3); }
}

int main(int argc, char** argv) {
    if (argc != 2) { std::cerr << "usage: mirror_driver <dir>\n"; return 2; }
    const std::string d = std::string(argv[1]) + "/";
    auto bases = read_all<G1Affine>(d + "bases.bin");
    auto scalars = read_all<BigInteger256>(d + "scalars.bin");
    auto r = VariableBase::msm(bases, scalars);
    check(r, "VariableBase::msm");
    write_all(d + "msm.out", &r.value, 1);

    {   // the generic arm of VariableBase::msm: Affine<G2> bases (optional input file)
        std::ifstream probe(d + "g2_bases.bin", std::ios::binary);
        if (probe.good()) {
            auto g2b = read_all<G2Affine>(d + "g2_bases.bin");
            auto g2s = read_all<BigInteger256>(d + "g2_scalars.bin");
            auto r2 = VariableBase::msm(g2b, g2s);
            check(r2, "VariableBase::msm (G2)");
            write_all(d + "msm_g2.out", &r2.value, 1);
        }
    }
    auto x = read_all<Fr>(d + "fr.bin");
    auto dom = EvaluationDomain::new_(x.size());
    if (!dom) return 2;
    const char* names[4] = {"fft.out", "ifft.out", "coset_fft.out", "coset_ifft.out"};
    for (int m = 0; m < 4; m++) {
        auto y = x;
        Result<Unit> s = m == 0 ? dom->fft_in_place(y) : m == 1 ? dom->ifft_in_place(y) : m == 2 ? dom->coset_fft_in_place(y) : dom->coset_ifft_in_place(y);
        check(s, names[m]);
        write_all(d + names[m], y.data(), y.size());
    }
    PolyMultiplier pm;
    pm.add_polynomial(read_all<Fr>(d + "p1.bin"), "p1");
    pm.add_polynomial(read_all<Fr>(d + "p2.bin"), "p2");
    auto prod = pm.multiply();
    if (!prod) return 2;
    check(*prod, "PolyMultiplier::multiply");
    write_all(d + "polymul.out", prod->value.data(), prod->value.size());
    bool threw = false;
    try { std::vector<Fr> bad(3); cuda::NTT(3, bad.data(), cuda::NTTInputOutputOrder::NN, cuda::NTTDirection::Forward, cuda::NTTType::Standard); }
    catch (const std::invalid_argument&) { threw = true; }     // the Rust shim panics here (lib.rs:84-86)
    std::printf("ok%s\n", threw ? "" : " (missing panic)");
    return threw ? 0 : 2;
}
