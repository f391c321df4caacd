"""Times the precomputed-table MSM against the windowed MSM on resident inputs (run on the GPU box)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from snarkvm_b200 import device

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for lg in [int(a) for a in sys.argv[1:]] or [20, 22, 24]:
    n = 1 << lg
    bases = device.generate_bases(n, seed=lg)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    scal = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    scal[:, 3] &= (1 << 60) - 1          # < r
    t0 = time.time(); pre = device.PrecomputedBases(bases); torch.cuda.synchronize(); tp = time.time() - t0
    a = device.msm(bases, scal); b = pre.msm(scal)
    print(f"lg={lg} equal={bool((a == b).all())} c={pre.c} nwin={pre.nwin} table={pre.table_bytes/2**30:.2f} GiB precompute={tp:.2f}s "
          f"windowed={t(lambda: device.msm(bases, scal)):.2f} ms precomputed={t(lambda: pre.msm(scal)):.2f} ms", flush=True)
    pre.free()
