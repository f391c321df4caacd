"""Focused plan sweep after a pair-level change (run on the GPU box): python tools/ab_pair.py [default-only]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from snarkvm_b200 import device

def run(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

default_only = len(sys.argv) > 1
grid = {24: ([16, 17, 18], [3, 4, 5, 6]), 22: ([15, 16, 17], [2, 3, 4, 5]), 20: ([13, 14, 15, 16], [0, 1, 2, 3]),
        18: ([11, 12, 13, 14], [0, 1, 2]), 16: ([10, 11, 12], [0, 1, 2])}
pre_grid = {24: ([22], [3, 4]), 22: ([20], [2, 3]), 20: ([17], [1, 2])}
rng = np.random.default_rng(0)
for lg in (16, 18, 20, 22, 24):
    n = 1 << lg
    bases = device.generate_bases(n, 7)
    s = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); s[:, 3] &= np.uint64((1 << 60) - 1)
    scal = torch.from_numpy(s.view(np.int64)).cuda()
    for k in ("SNARKVM_B200_MSM_C", "SNARKVM_B200_MSM_LEVELS"): os.environ.pop(k, None)
    ref = device.msm(bases, scal)
    print(f"lg={lg} default plan: {run(lambda: device.msm(bases, scal)):.2f} ms", flush=True)
    if default_only: continue
    cs, ls = grid[lg]
    for c in cs:
        row = []
        for L in ls:
            os.environ["SNARKVM_B200_MSM_C"] = str(c); os.environ["SNARKVM_B200_MSM_LEVELS"] = str(L)
            ok = bool((device.msm(bases, scal) == ref).all())
            row.append(f"L={L}: {run(lambda: device.msm(bases, scal)):6.2f}{'' if ok else ' MISMATCH'}")
        print(f"lg={lg} c={c}  " + "  ".join(row), flush=True)
    for k in ("SNARKVM_B200_MSM_C", "SNARKVM_B200_MSM_LEVELS"): os.environ.pop(k, None)
    if lg in pre_grid:
        cs, ls = pre_grid[lg]
        for c in cs:
            for L in ls:
                os.environ["SNARKVM_B200_MSM_PRE_C"] = str(c); os.environ["SNARKVM_B200_MSM_PRE_LEVELS"] = str(L)
                pre = device.PrecomputedBases(bases)
                ok = bool((pre.msm(scal) == ref).all())
                print(f"lg={lg} precomputed c={c} L={L}: {run(lambda: pre.msm(scal)):.2f} ms{'' if ok else ' MISMATCH'}", flush=True)
                pre.free()
