#!/usr/bin/env python
"""bench.py — BLS12-377 G1 MSM points/s (headline) and Fr NTT elements/s (secondary) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference CPU algorithms (oracle port)

One "step" = one pass of the hot path over one batch of synthetic input:
  * product arm: one 2^lg-point VariableBase::msm per GPU (weak scaling: every rank owns 2^lg points;
    at N > 1 the per-rank window sums are all-gathered over NCCL and added on the device) — `value` is
    measured with bases and scalars resident in HBM, `e2e` through the reference-facing C-ABI symbol
    `snarkvm_msm` with pinned HOST buffers (H2D of points + scalars and D2H of the result inside the
    timed region).
  * reference arm: the CPU restatement of batched::msm (oracle/oracle.c, OpenMP over all host cores) on a
    bounded sample of the same workload — the Rust reference cannot be built here (no cargo/rustc).

Prints ONE JSON line on rank 0.  Timing: CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks; inputs (≥ 2.2 GB per step) are far larger than L2 so no flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_POINT = 128      # 96 B affine (x, y) + 32 B scalar, each touched once
ALGO_BYTES_PER_ELEMENT = 64     # NTT: 32 B read + 32 B write per element per transform
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def load_traffic(key):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full summary, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def random_scalars(n, seed):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 60) - 1)          # < r (r ≈ 2^252.1): uniform below 2^252
    return s


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (CPU restatement of the reference algorithms)
# ------------------------------------------------------------------------------------------------
def cpu_msm_points_per_s(lg_sample, bases_host, reps=1):
    from oracle import cpu
    n = 1 << lg_sample
    scal = random_scalars(n, 777)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        cpu.msm(bases_host[:n], scal, cpu.BATCHED)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return n / best, best


def cpu_ntt_elements_per_s(lg_sample):
    """Best over a few OpenMP team sizes: the port's per-stage parallel-for (standing in for rayon's chunked
    butterflies, fft/domain.rs:667-688) stops scaling — and can slow down — on very wide hosts."""
    from oracle import cpu
    x = random_scalars(1 << lg_sample, 778)
    allc = cpu.num_threads()
    best = None
    for th in sorted({allc, min(allc, 64), min(allc, 32), min(allc, 16)}, reverse=True):
        cpu.set_num_threads(th)
        cpu.ntt(x[:1 << 12], cpu.FORWARD, cpu.STANDARD)          # spin the team up
        t = time.perf_counter()
        cpu.ntt(x, cpu.FORWARD, cpu.STANDARD)
        dt = time.perf_counter() - t
        if best is None or dt < best[0]:
            best = (dt, th)
    cpu.set_num_threads(allc)
    return (1 << lg_sample) / best[0], best[0], best[1]


def cpu_bases(n):
    """Valid subgroup points for the CPU legs, built with the oracle's batched affine addition."""
    from oracle import bls12_377 as py, cpu
    bases = np.zeros((n, 104), dtype=np.uint8)
    bases[0] = np.frombuffer(py.affine_bytes(py.g1_mul(py.G1_GENERATOR, 0xB200)), dtype=np.uint8)
    cur = 1
    while cur < n:
        m = min(cur, n - cur)
        step = np.frombuffer(py.affine_bytes(py.g1_mul(py.G1_GENERATOR, cur)), dtype=np.uint8)
        bases[cur:cur + m] = cpu.batch_affine_add(bases[:m], np.tile(step, (m, 1)))
        cur += m
    return bases


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import cpu
    cpu.build()
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it may run on (rayon's default)
    try:
        cpu.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        cpu.set_num_threads(os.cpu_count() or 1)
    lg = args.ref_lg
    bases = cpu_bases(1 << lg)
    threads = cpu.num_threads()
    for _ in range(args.warmup):
        cpu_msm_points_per_s(min(lg, 16), bases)
    scal = random_scalars(1 << lg, 777)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.msm(bases, scal, cpu.BATCHED)
    dt = (time.perf_counter() - t0) / args.steps
    value = (1 << lg) / dt
    ntt_v, ntt_dt, ntt_th = cpu_ntt_elements_per_s(args.ref_ntt_lg)
    sample = f"2^{lg}-point batched::msm per step on {threads} OpenMP threads (one task per window, as rayon in batched.rs:400-401)"
    line = {
        "impl": "reference", "metric": "bls12_377_g1_msm_points_per_sec", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (CPU)", "data": "synthetic",
        "config": {"workload": f"bls12_377_g1_msm_2^{args.lg}_points_per_gpu", "sample": sample,
                   "note": "C restatement of the reference CPU (rayon) path — Rust toolchain unavailable"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ntt": {"metric": "fr_ntt_elements_per_sec", "value": ntt_v, "unit": "elements/s",
                "sample": f"one 2^{args.ref_ntt_lg} forward fft_in_place, best of 16/32/64/{threads} threads = {ntt_th}", "ms": ntt_dt * 1e3},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def run_product(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from snarkvm_b200 import _lib, cuda as shim, device, sharded

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = 1 << args.lg
    peak, peak_src = load_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs: per-rank shard, generated in HBM; host copies in pinned memory for the e2e leg ----
    bases = device.generate_bases(n, seed=0xB200 + rank, device=dev)
    scal_host = torch.from_numpy(random_scalars(n, 1234 + rank).view(np.int64)).pin_memory()
    scalars = scal_host.to(dev)
    torch.cuda.synchronize()

    def step_device():
        return sharded.msm_sharded(bases, scalars) if world > 1 else device.msm(bases, scalars)

    # ---- correctness guard on the timed configuration (closed form: bases are known multiples of G) ----
    first = step_device()

    # ---- `value`: device-resident MSM, K steps, CUDA events on the launching stream ----
    for _ in range(args.warmup):
        step_device()
    _lib.profile_enable(True)
    for k in (0, 1, 2, 3):
        _lib.profile_collect(k)
    sampler = ClockSampler(local_rank)
    launches0 = _lib.launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.launch_count() - launches0
    acc_ms, acc_n = _lib.profile_collect(_lib.PROF_MSM_ACCUMULATE)
    sort_ms, _ = _lib.profile_collect(_lib.PROF_MSM_SORT)
    red_ms, _ = _lib.profile_collect(_lib.PROF_MSM_REDUCE)
    _lib.profile_enable(False)
    assert (out == first).all(), "MSM result changed between steps"
    ms_per_step = ms_total / args.steps
    value = n * world / (ms_per_step * 1e-3)
    acc_ms_per_launch = acc_ms / max(1, args.steps)        # the whole accumulation phase of one MSM (several kernels)
    achieved = ALGO_BYTES_PER_POINT * n / (acc_ms_per_launch * 1e-3) / 1e9

    # ---- `e2e`: through the drop-in C-ABI symbol with pinned HOST buffers (H2D + D2H inside the timing) ----
    bases_host = torch.empty((n, 104), dtype=torch.uint8).pin_memory()
    bases_host.copy_(bases)
    torch.cuda.synchronize()
    b_np, s_np = bases_host.numpy(), scal_host.numpy().view(np.uint64)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def step_e2e():
        if world == 1:
            return shim.msm(b_np, s_np)                     # snarkvm_msm: H2D points+scalars, MSM, D2H 144 B
        db = bases_host.to(dev, non_blocking=True)           # sharded public API: same copies, then the NCCL exchange
        ds = scal_host.to(dev, non_blocking=True)
        return sharded.msm_sharded(db, ds)

    r_e2e = step_e2e()
    assert (r_e2e == first).all(), "e2e result differs from the device-resident result"
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e_value = n * world / (e2e_ms * 1e-3)
    e2e_resident = None
    if world == 1:
        # same call, SRS-style usage: the bases slice was registered once (snarkvm_b200_register_bases), only the
        # 32 B/point scalars cross PCIe per call
        shim.register_bases(b_np)
        assert (shim.msm(b_np, s_np) == first).all()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            shim.msm(b_np, s_np)
        torch.cuda.synchronize()
        r_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        shim.unregister_bases(b_np)
        e2e_resident = {"value": n / (r_ms * 1e-3), "unit": "points/s", "ms_per_step": r_ms, "h2d_bytes_per_step": n * 32,
                        "d2h_bytes_per_step": 144, "api": "snarkvm_msm after snarkvm_b200_register_bases (bases resident)"}
        # and with the fixed-base tables of the registered slice built once (snarkvm_b200_register_bases_precomputed);
        # an extra, never the headline: a failure here (e.g. not enough HBM for the tables) is reported, not fatal
        try:
            t0 = time.perf_counter()
            shim.register_bases_precomputed(b_np)
            setup_s = time.perf_counter() - t0
            try:
                assert (shim.msm(b_np, s_np) == first).all()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    shim.msm(b_np, s_np)
                torch.cuda.synchronize()
                p_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
            finally:
                shim.unregister_bases(b_np)
            e2e_resident["precomputed_tables"] = {"value": n / (p_ms * 1e-3), "unit": "points/s", "ms_per_step": p_ms,
                                                  "one_time_setup_s": setup_s,
                                                  "api": "snarkvm_msm after snarkvm_b200_register_bases_precomputed"}
        except Exception as exc:  # noqa: BLE001
            e2e_resident["precomputed_tables"] = {"error": repr(exc)}

    # ---- secondary metric: Fr NTT elements/s (per-GPU replicas; BASELINE config 3) ----
    ntt = None
    if not args.skip_ntt:
        from snarkvm_b200.cuda import NTTDirection, NTTType
        nn = 1 << args.ntt_lg
        x = torch.from_numpy(random_scalars(nn, 99 + rank).view(np.int64)).to(dev)
        scratch = torch.empty_like(x)
        for _ in range(3):
            device.ntt_(x, NTTDirection.Forward, NTTType.Standard, scratch)
        _lib.profile_enable(True)
        _lib.profile_collect(_lib.PROF_NTT_PASS)
        barrier()
        reps = 10
        e0.record()
        for _ in range(reps):
            device.ntt_(x, NTTDirection.Forward, NTTType.Standard, scratch)
        e1.record()
        barrier()
        ntt_ms = max_over_ranks(e0.elapsed_time(e1)) / reps
        pass_ms, pass_n = _lib.profile_collect(_lib.PROF_NTT_PASS)
        _lib.profile_enable(False)
        x_h = torch.from_numpy(random_scalars(nn, 5).view(np.int64)).pin_memory()
        xs = x_h.numpy().view(np.uint64)
        shim.NTT(nn, xs, shim.NTTInputOutputOrder.NN, shim.NTTDirection.Forward, shim.NTTType.Standard)
        barrier()
        t0 = time.perf_counter()
        shim.NTT(nn, xs, shim.NTTInputOutputOrder.NN, shim.NTTDirection.Forward, shim.NTTType.Standard)
        barrier()
        ntt_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        per_pass_ms = pass_ms / max(1, pass_n)
        ntt = {
            "metric": "fr_ntt_elements_per_sec", "value": nn * world / (ntt_ms * 1e-3), "unit": "elements/s",
            "workload": f"forward NN NTT of 2^{args.ntt_lg} Fr elements per GPU (replicas, no collective)",
            "ms_per_transform": ntt_ms, "passes_per_transform": pass_n // reps if reps else None,
            "e2e": {"value": nn * world / (ntt_e2e_ms * 1e-3), "unit": "elements/s", "h2d_bytes_per_step": nn * 32,
                    "d2h_bytes_per_step": nn * 32, "api": "snarkvm_ntt (host buffer, pinned)"},
            "roofline": {"bound": "hbm", "kernel": "k_ntt_pass", "achieved": ALGO_BYTES_PER_ELEMENT * nn / (per_pass_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": ALGO_BYTES_PER_ELEMENT * nn / (per_pass_ms * 1e-3) / 1e9 / peak,
                         "traffic": load_traffic("k_ntt_pass"), "per_launch_ms": per_pass_ms,
                         "note": "per pass: 64 B/element algorithmic; a transform is %d passes" % (pass_n // reps)},
        }

    # ---- BASELINE config 4: KZG10 commit core at 2^22 coefficients (Montgomery → canonical → MSM), operands resident ----
    kzg = None
    if not args.skip_kzg:
        nk = 1 << args.kzg_lg
        coeffs = torch.from_numpy(random_scalars(nk, 4242 + rank).view(np.int64)).to(dev)
        powers = bases[:nk] if nk <= n else device.generate_bases(nk, seed=0xB200 + rank, device=dev)
        for _ in range(3):
            device.kzg_commit(powers, coeffs)
        barrier()
        e0.record()
        for _ in range(5):
            device.kzg_commit(powers, coeffs)
        e1.record()
        barrier()
        kzg_ms = max_over_ranks(e0.elapsed_time(e1)) / 5
        kzg = {"metric": "kzg10_commit_coefficients_per_sec", "value": nk * world / (kzg_ms * 1e-3), "unit": "coefficients/s",
               "ms_per_commit": kzg_ms, "workload": f"2^{args.kzg_lg}-coefficient polynomial, powers resident in HBM, per GPU"}
        # same commitment over precomputed tables 2^{c·w}·P_i of the powers (one bucket set for all windows)
        t0 = time.perf_counter()
        pre = device.PrecomputedBases(powers)
        torch.cuda.synchronize()
        pre_s = time.perf_counter() - t0
        assert (pre.kzg_commit(coeffs) == device.kzg_commit(powers, coeffs)).all()
        for _ in range(2):
            pre.kzg_commit(coeffs)
        barrier()
        e0.record()
        for _ in range(5):
            pre.kzg_commit(coeffs)
        e1.record()
        barrier()
        pre_ms = max_over_ranks(e0.elapsed_time(e1)) / 5
        kzg["precomputed_tables"] = {"value": nk * world / (pre_ms * 1e-3), "unit": "coefficients/s", "ms_per_commit": pre_ms,
                                     "window_bits": pre.c, "windows": pre.nwin, "table_bytes": pre.table_bytes,
                                     "one_time_precompute_s": pre_s}
        pre.free()
        # UniversalParams::lagrange_basis: iFFT over G1 points (data_structures.rs:68-72)
        ng = 1 << args.g1_ntt_lg
        gp = powers.reshape(-1)[: ng * 104].contiguous()
        device.lagrange_basis(gp)
        barrier()
        e0.record()
        device.lagrange_basis(gp)
        e1.record()
        barrier()
        kzg["lagrange_basis"] = {"ms": max_over_ranks(e0.elapsed_time(e1)), "workload": f"ifft of 2^{args.g1_ntt_lg} G1 points per GPU",
                                 "unit": "ms"}

    if rank != 0:
        return

    # ---- cpu_baseline: the oracle port on this box's host cores, bounded sample (N = 1 only) ----
    cpu_baseline = None
    if world == 1 and not args.skip_cpu:
        from oracle import cpu
        cpu.build()
        try:
            cpu.set_num_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            cpu.set_num_threads(os.cpu_count() or 1)
        lg_s = args.cpu_lg
        hb = b_np[: 1 << lg_s]
        v, dt = cpu_msm_points_per_s(lg_s, hb)
        cpu_baseline = {"value": v, "unit": "points/s", "cores": cpu.num_threads(), "kind": "port",
                        "sample": f"one 2^{lg_s}-point batched::msm ({dt:.1f} s) — C restatement of the reference CPU path, "
                                  f"OpenMP task per window"}
        if ntt is not None:
            nv, ndt, nth = cpu_ntt_elements_per_s(args.cpu_ntt_lg)
            ntt["cpu_baseline"] = {"value": nv, "unit": "elements/s", "cores": nth, "kind": "port",
                                   "sample": f"one 2^{args.cpu_ntt_lg} forward fft_in_place ({ndt:.2f} s), best team size of 16/32/64/all"}

    plan = device.msm_plan(n)
    plan_levels = 4 if args.lg >= 21 else 1 if args.lg == 20 else 0      # msm_make_plan (csrc/msm.cu)
    line = {
        "metric": "bls12_377_g1_msm_points_per_sec", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (Montgomery integer arithmetic, IMAD pipe)",
        "data": "synthetic",
        "config": {"workload": f"bls12_377_g1_msm_2^{args.lg}_points_per_gpu", "points_per_gpu": n, "scalars": "uniform < 2^252",
                   "bases": "h(seed,i)*G, affine 104 B stride", "window_bits": plan["c"], "windows": plan["nwin"],
                   "parallelism": f"points sharded over {world} GPU(s); one all-gather of window sums" if world > 1 else "single GPU",
                   "l2": f"inputs ({n * 136 / 1e9:.2f} GB/step) plus {n * plan['nwin'] * 4 / 1e9:.2f} GB of sorted entries and the dense pair-level scratch stream through the 126 MB L2 every step — no flush needed"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": n * (104 + 32) * world,
                "d2h_bytes_per_step": 144 * world, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "api": "snarkvm_msm (drop-in C-ABI, pinned host buffers)" if world == 1 else "sharded.msm_sharded after H2D"},
        "e2e_registered_bases": e2e_resident,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "bucket accumulation phase: k_densify_bases + k_pair_level x%d + k_bucket_accumulate%s" % (
                         plan_levels, "_dense" if plan_levels else ""),
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": load_traffic("msm_bucket_accumulation_phase") if args.lg == 24 else None,
                     "peak_source": peak_src, "per_launch_ms": acc_ms_per_launch,
                     "share_of_step": {"sort": sort_ms / ms_total, "accumulate": acc_ms / ms_total, "reduce": red_ms / ms_total},
                     "note": "algorithmic 128 B/point over the phase's duration; the phase is bound by the INT32 multiplier (fmaheavy pipe "
                             "65-87 % active, profiles/r1_msm_metrics.csv), not by HBM"},
        "cpu_baseline": cpu_baseline,
        "ntt": ntt,
        "kzg_commit": kzg,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lg", type=int, default=24, help="log2 points per GPU (BASELINE configs[1]: 2^20/2^22/2^24)")
    ap.add_argument("--ntt-lg", type=int, default=24)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-lg", type=int, default=21, help="cpu_baseline sample size (bounded)")
    ap.add_argument("--cpu-ntt-lg", type=int, default=22)
    ap.add_argument("--ref-lg", type=int, default=20, help="--impl reference: points per step")
    ap.add_argument("--ref-ntt-lg", type=int, default=22)
    ap.add_argument("--kzg-lg", type=int, default=22)
    ap.add_argument("--g1-ntt-lg", type=int, default=12, help="size of the G1 iFFT (lagrange_basis) timed inside the KZG extra")
    ap.add_argument("--skip-kzg", action="store_true")
    ap.add_argument("--skip-ntt", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_product(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
