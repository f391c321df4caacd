"""Sweep window bits × pair levels for the precomputed-table MSM (run on the GPU box).  usage: tune_precomputed.py lg c1,c2 l1,l2"""
import ctypes, os, sys
import torch
sys.path.insert(0, ".")
from snarkvm_b200 import device, _lib

def prof(fn, reps=3):
    L = _lib.lib()
    fn(); torch.cuda.synchronize()
    L.snarkvm_b200_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    out = []
    for k in range(3):
        ms, cnt = ctypes.c_double(), ctypes.c_uint64()
        L.snarkvm_b200_profile_collect(k, ctypes.byref(ms), ctypes.byref(cnt))
        out.append(ms.value / reps)
    L.snarkvm_b200_profile_enable(0)
    return e0.elapsed_time(e1) / reps, out

lg = int(sys.argv[1]); n = 1 << lg
cs = [int(x) for x in sys.argv[2].split(",")]
ls = [int(x) for x in sys.argv[3].split(",")]
bases = device.generate_bases(n, seed=lg)
g = torch.Generator(device="cuda"); g.manual_seed(lg)
scal = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
scal[:, 3] &= (1 << 60) - 1
ref = device.msm(bases, scal)
for c in cs:
    os.environ["SNARKVM_B200_MSM_PRE_C"] = str(c)
    pre = device.PrecomputedBases(bases)
    for l in ls:
        os.environ["SNARKVM_B200_MSM_PRE_LEVELS"] = str(l)
        # the plan is fixed in the handle at precompute time except for levels… rebuild to be safe
        pre.free(); pre = device.PrecomputedBases(bases)
        ok = bool((pre.msm(scal) == ref).all())
        tot, (s, a, r) = prof(lambda: pre.msm(scal))
        print(f"lg={lg} c={c} nwin={pre.nwin} levels={l} ok={ok} total={tot:.2f} sort={s:.2f} acc={a:.2f} reduce={r:.2f}", flush=True)
    pre.free()
