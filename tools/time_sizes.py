"""Default-plan MSM time (inputs resident) at the given log2 sizes: python tools/time_sizes.py 14 23 25 26   (run on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from snarkvm_b200 import device

for lg in [int(a) for a in sys.argv[1:]]:
    n = 1 << lg
    bases = device.generate_bases(n, 7)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    scal = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    scal[:, 3] &= (1 << 60) - 1
    device.msm(bases, scal); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if lg >= 24 else 10
    e0.record()
    for _ in range(reps): device.msm(bases, scal)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"lg={lg} plan={device.msm_plan(n)} {ms:.2f} ms  {n / ms / 1e3:.1f} Mpoints/s", flush=True)
    del bases, scal
    torch.cuda.empty_cache()
