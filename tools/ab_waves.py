"""Default-plan MSM time per size and a sweep of the pair-level wave count (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from snarkvm_b200 import device

def run(fn, reps=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

rng = np.random.default_rng(0)
waves = [int(a) for a in sys.argv[1:]]
for lg in (20, 22, 24):
    n = 1 << lg
    bases = device.generate_bases(n, 7)
    s = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); s[:, 3] &= np.uint64((1 << 60) - 1)
    scal = torch.from_numpy(s.view(np.int64)).cuda()
    os.environ.pop("SNARKVM_B200_MSM_PAIR_WAVES", None)
    row = [f"default {run(lambda: device.msm(bases, scal)):.2f}"]
    for w in waves:
        os.environ["SNARKVM_B200_MSM_PAIR_WAVES"] = str(w)
        row.append(f"waves={w} {run(lambda: device.msm(bases, scal)):.2f}")
    print(f"lg={lg}  " + "  ".join(row), flush=True)
