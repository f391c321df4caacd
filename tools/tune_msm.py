"""Sweep MSM plan parameters (window bits c, pair levels) per size on the GPU; prints ms per MSM.
    python tools/tune_msm.py [lg ...]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from snarkvm_b200 import device

def run(bases, scal, reps=3):
    device.msm(bases, scal)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        device.msm(bases, scal)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

lgs = [int(a) for a in sys.argv[1:]] or [16, 18, 20, 22, 24]
rng = np.random.default_rng(0)
for lg in lgs:
    n = 1 << lg
    bases = device.generate_bases(n, 7)
    s = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); s[:, 3] &= np.uint64((1 << 60) - 1)
    scal = torch.from_numpy(s.view(np.int64)).cuda()
    best = None
    for c in range(max(8, lg - 9), min(21, lg - 2) + 1):
        row = []
        for L in (0, 1, 2, 3, 4, 6, 8):
            avg = n >> (c - 1)
            if L and (avg >> L) < 2: row.append("   -  "); continue
            os.environ["SNARKVM_B200_MSM_C"] = str(c); os.environ["SNARKVM_B200_MSM_LEVELS"] = str(L)
            try:
                ms = run(bases, scal)
            except Exception as e:
                row.append(" err  "); continue
            row.append(f"{ms:6.2f}")
            if best is None or ms < best[0]: best = (ms, c, L)
        print(f"lg={lg} c={c:2d} W={253 // c + 1:2d} L=0,1,2,3,4,6,8: " + " ".join(row), flush=True)
    print(f"lg={lg} BEST {best[0]:.2f} ms c={best[1]} L={best[2]}  -> {n / best[0] / 1e3:.1f} Mpts/s", flush=True)
