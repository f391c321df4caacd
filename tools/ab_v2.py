"""A/B of the pair-level kernels (round-1 thread-contiguous v1 vs warp-interleaved smem-ring v2) with the per-phase split,
plus the scratch-group budget and the small sizes: python tools/ab_v2.py   (run on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from snarkvm_b200 import _lib, device

def run(fn, reps):
    fn(); torch.cuda.synchronize()
    _lib.profile_enable(True)
    for k in range(4): _lib.profile_collect(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ph = [_lib.profile_collect(k)[0] / reps for k in range(3)]
    _lib.profile_enable(False)
    return e0.elapsed_time(e1) / reps, ph

rng = np.random.default_rng(0)
sizes = [int(a) for a in sys.argv[1:]] or [24, 22, 20]
for lg in sizes:
    n = 1 << lg
    bases = device.generate_bases(n, 7)
    s = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); s[:, 3] &= np.uint64((1 << 60) - 1)
    scal = torch.from_numpy(s.view(np.int64)).cuda()
    ref = None
    KEYS = ("SNARKVM_B200_MSM_PAIR_V1", "SNARKVM_B200_MSM_SCRATCH_GB", "SNARKVM_B200_MSM_LEVELS", "SNARKVM_B200_MSM_C", "SNARKVM_B200_MSM_PAIR_PARTS", "SNARKVM_B200_MSM_PAIR_MINB")
    KEYS = KEYS + ("SNARKVM_B200_MSM_RECORDS",)
    if lg >= 20:
        variants = (("default", {}), ("gather (r1 sort)", {"SNARKVM_B200_MSM_RECORDS": "0"}))
    else:
        variants = (("default", {}), ("L0 c16", {"SNARKVM_B200_MSM_LEVELS": "0", "SNARKVM_B200_MSM_C": "16"}))
    for tag, env in variants:
        for k in KEYS: os.environ.pop(k, None)
        os.environ.update(env)
        got = device.msm(bases, scal)
        if ref is None: ref = got
        ms, ph = run(lambda: device.msm(bases, scal), 3 if lg >= 24 else 8)
        print(f"lg={lg} {tag:16s} {ms:8.2f} ms  sort {ph[0]:6.2f}  accumulate {ph[1]:7.2f}  reduce {ph[2]:6.2f}  {'ok' if (got == ref).all() else 'MISMATCH'}", flush=True)
    for k in KEYS: os.environ.pop(k, None)
    del bases, scal
    torch.cuda.empty_cache()
for lg in (10, 12, 14, 16, 18):
    n = 1 << lg
    bases = device.generate_bases(n, 7)
    s = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); s[:, 3] &= np.uint64((1 << 60) - 1)
    scal = torch.from_numpy(s.view(np.int64)).cuda()
    ref = None
    for tag, env in (("thread path", {"SNARKVM_B200_MSM_WARP_PATH": "0"}), ("warp path", {})):
        os.environ.pop("SNARKVM_B200_MSM_WARP_PATH", None)
        os.environ.update(env)
        got = device.msm(bases, scal)
        if ref is None: ref = got
        ms, ph = run(lambda: device.msm(bases, scal), 20)
        print(f"lg={lg} {tag:12s} {ms:8.3f} ms  sort {ph[0]:6.3f}  accumulate {ph[1]:7.3f}  reduce {ph[2]:6.3f}  {'ok' if (got == ref).all() else 'MISMATCH'}", flush=True)
    os.environ.pop("SNARKVM_B200_MSM_WARP_PATH", None)
print("scratch:", device.msm_scratch_stats())
