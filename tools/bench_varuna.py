"""Varuna prover rounds + round commitments on a synthetic R1CS with every polynomial resident in HBM (SURVEY §8 f3,
BASELINE configs[4] "Varuna prove() on a 2^20-constraint synthetic R1CS"):   python tools/bench_varuna.py [lg ...]

What one "proof" is here: init_prover (z_A, z_B, z_C), the five AHP rounds (w; h_0; g_1, h_1; g_a, g_b, g_c; h_2) of the
reference's TestCircuit (one instance, non-hiding mode) and the KZG commitment of every round's oracles as ONE batched MSM pass
per round over resident powers (sonic_pc/mod.rs:177-257).  Not included: the Fiat-Shamir sponge (Poseidon, host side —
challenges are fixed numbers here) and the final batch opening.  Prints one JSON line per size."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from snarkvm_b200 import device, varuna
from snarkvm_b200.algorithms import KZG10


def run(lg, reps=2, mul_depth=3):
    dev = torch.device("cuda:0")
    n = 1 << lg
    circuit, z = varuna.test_circuit_csr(3, 5, mul_depth, n, n, dev)
    powers = device.generate_bases(4 * n, seed=2024, device=dev)              # synthetic SRS-shaped bases (no trapdoor needed for timing)
    alpha, eta_b, eta_c, beta = 0x1234567, 0x2345678, 0x3456789, 0x456789A
    deltas = [1, 0x56789AB, 0x6789ABC]
    out = {}

    def once():
        t = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p = varuna.Prover(circuit, [z])
        torch.cuda.synchronize(); t["init"] = time.perf_counter() - t0
        comms = []
        def stage(name, fn, polys_of):
            torch.cuda.synchronize(); a = time.perf_counter()
            r = fn()
            torch.cuda.synchronize(); b = time.perf_counter()
            polys = [q.contiguous() for q in polys_of(r) if q.shape[0]]
            assert all(q.shape[0] <= powers.shape[0] for q in polys)
            comms.append(KZG10.batch_commit(powers, polys))
            torch.cuda.synchronize(); c = time.perf_counter()
            t[name] = b - a; t[name + "_commit"] = c - b
        stage("round1", p.first_round, lambda r: r)
        p.assignments()
        stage("round2", p.second_round, lambda r: [r])
        stage("round3", lambda: p.third_round(alpha, eta_b, eta_c), lambda r: list(r))
        stage("round4", lambda: p.fourth_round(alpha, beta), lambda r: list(r))
        stage("round5", lambda: p.fifth_round(deltas), lambda r: [r])
        torch.cuda.synchronize(); t["total"] = time.perf_counter() - t0
        return t, comms
    once()                                                                      # warm-up: twiddle tables, pools
    best = None
    for _ in range(reps):
        t, comms = once()
        if best is None or t["total"] < best["total"]: best = t
    out = {"metric": "varuna_prover_rounds_constraints_per_sec", "value": n / best["total"], "unit": "constraints/s",
           "constraints": n, "variables": n, "s_per_proof": best["total"],
           "phases_ms": {k: round(v * 1e3, 2) for k, v in best.items() if k != "total"},
           "workload": f"TestCircuit 2^{lg} constraints × 2^{lg} variables, 1 instance, non-hiding; rounds 1–5 + one batched KZG commit pass per round; "
                       "sponge and batch opening not included"}
    del powers
    torch.cuda.empty_cache()
    return out


def run_prove(lg, reps=2, mul_depth=3, hiding=True):
    """The whole prove_batch-shaped pipeline (varuna.rs:336-620) for one instance of the TestCircuit: init, rounds 1-5 with the hiding
    mode's mask polynomial, SonicKZG10::commit of every round with the reference's degree and hiding bounds (one device pass per
    round), construct_linear_combinations, the evaluations the proof carries, and open_combinations (three opening proofs).  Only the
    Poseidon sponge is missing: the challenges are fixed numbers."""
    from snarkvm_b200.sonic_pc import CommitterKey, LabeledPolynomial, SonicKZG10
    dev = torch.device("cuda:0")
    n = 1 << lg
    circuit, z = varuna.test_circuit_csr(3, 5, mul_depth, n, n, dev)
    D = 4 * n + 8
    powers = device.generate_bases(D + 1, seed=2024, device=dev)                # SRS-shaped bases (timing needs no trapdoor)
    gpowers = device.generate_bases(D + 2, seed=2025, device=dev)
    V, K = circuit.variable_domain, circuit.max_non_zero_domain
    bounds = {"g_1": V.size - 2, "g_a": circuit.ariths[0].domain.size - 2, "g_b": circuit.ariths[1].domain.size - 2, "g_c": circuit.ariths[2].domain.size - 2}
    ck = CommitterKey.trim(powers, gpowers, supported_degree=D, supported_hiding_bound=1, enforced_degree_bounds=sorted(set(bounds.values())))
    alpha, eta_b, eta_c, beta, gamma = 0x1234567, 0x2345678, 0x3456789, 0x456789A, 0x56789AB
    deltas = [1, 0x56789AB, 0x6789ABC]
    blind3 = torch.from_numpy(__import__("numpy").array([varuna._mont(v) for v in (3, 5, 7)], dtype="uint64").view("int64")).to(dev)

    def once():
        t = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p = varuna.Prover(circuit, [z])
        if hiding:
            p.set_mask_poly([11, 12, 13, 14], [0, 21, 22, 23, 24, 25])
        torch.cuda.synchronize(); t["init"] = time.perf_counter() - t0
        labeled, rands = {}, {}
        def stage(name, fn, names):
            torch.cuda.synchronize(); a = time.perf_counter()
            fn()
            torch.cuda.synchronize(); b = time.perf_counter()
            polys = p_polys()
            lps, bl = [], []
            for k in names:
                if k not in polys:
                    continue
                q = polys[k]
                if k in bounds and q.shape[0] > bounds[k] + 1:
                    q = q[: bounds[k] + 1]
                hide = hiding and (k.startswith("w_") or k in bounds)
                lps.append(LabeledPolynomial(k, q.contiguous(), bounds.get(k), 1 if hide else None)); bl.append(blind3 if hide else None)
            _, rs = SonicKZG10.commit(ck, lps, bl)
            for lp, r in zip(lps, rs):
                labeled[lp.label] = lp; rands[lp.label] = r
            torch.cuda.synchronize(); c = time.perf_counter()
            t[name] = b - a; t[name + "_commit"] = c - b
        def p_polys():
            out = {f"w_{j}": w for j, w in enumerate(p.w_polys)}
            if p.mask_poly is not None: out["mask_poly"] = p.mask_poly
            for k in ("h_0", "g_1", "h_1", "h_2"):
                if getattr(p, k, None) is not None: out[k] = getattr(p, k)
            for m, g in zip("abc", getattr(p, "gs", [])): out[f"g_{m}"] = g
            return out
        stage("round1", p.first_round, ["w_0", "mask_poly"])
        p.assignments()
        stage("round2", p.second_round, ["h_0"])
        stage("round3", lambda: p.third_round(alpha, eta_b, eta_c), ["g_1", "h_1"])
        stage("round4", lambda: p.fourth_round(alpha, beta), ["g_a", "g_b", "g_c"])
        stage("round5", lambda: p.fifth_round(deltas), ["h_2"])
        torch.cuda.synchronize(); a = time.perf_counter()
        lcs, qs = p.linear_combinations(alpha, eta_b, eta_c, beta, deltas, gamma)
        polys = p.polynomials()
        for k in ("a_poly_a", "a_poly_b", "a_poly_c", "b_poly_a", "b_poly_b", "b_poly_c"):      # index polynomials: not committed by the prover
            labeled[k] = LabeledPolynomial(k, polys[k].contiguous(), None, None); rands[k] = type(rands["h_0"])()
        evals = {k: p._eval(labeled[k].polynomial, pt) for k, (_, pt) in qs if k in ("g_1", "g_a", "g_b", "g_c")}      # proof::Evaluations
        torch.cuda.synchronize(); b = time.perf_counter()
        names = sorted(labeled)
        proofs = SonicKZG10.open_combinations(ck, lcs, [labeled[k] for k in names], [rands[k] for k in names], qs, iter(range(1000, 1100)))
        torch.cuda.synchronize(); c = time.perf_counter()
        t["linear_combinations"] = b - a; t["open_combinations"] = c - b
        t["total"] = time.perf_counter() - t0
        assert len(proofs) == 3 and len(evals) == 4
        return t
    once()
    best = None
    for _ in range(reps):
        t = once()
        if best is None or t["total"] < best["total"]: best = t
    out = {"metric": "varuna_prove_constraints_per_sec", "value": n / best["total"], "unit": "constraints/s",
           "constraints": n, "variables": n, "s_per_proof": best["total"],
           "phases_ms": {k: round(v * 1e3, 2) for k, v in best.items() if k != "total"},
           "workload": f"TestCircuit 2^{lg} constraints × 2^{lg} variables, 1 instance, {'hiding' if hiding else 'non-hiding'} mode; init + rounds 1–5 + "
                       "SonicKZG10::commit per round (degree and hiding bounds) + linear combinations + open_combinations; Poseidon sponge not included"}
    del powers, gpowers
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "prove":
        for lg in [int(a) for a in sys.argv[2:]] or [16, 18, 20]:
            print(json.dumps(run_prove(lg)), flush=True)
        sys.exit(0)
    for lg in [int(a) for a in sys.argv[1:]] or [16, 18, 20]:
        print(json.dumps(run(lg)), flush=True)
