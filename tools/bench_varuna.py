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


if __name__ == "__main__":
    for lg in [int(a) for a in sys.argv[1:]] or [16, 18, 20]:
        print(json.dumps(run(lg)), flush=True)
