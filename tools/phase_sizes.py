"""Where a default-plan MSM spends its time at each size, warm: total (incl. D2H + host Horner), the device part alone
(window sums left in HBM) and the library's own event brackets (sort / accumulate / reduce).
python tools/phase_sizes.py 10 12 14 16 18 20   (run on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from snarkvm_b200 import device, _lib

for lg in [int(a) for a in sys.argv[1:]]:
    n = 1 << lg
    bases = device.generate_bases(n, 7)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    scal = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    scal[:, 3] &= (1 << 60) - 1
    reps = 50
    for _ in range(20): device.msm(bases, scal)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): device.msm(bases, scal)
    total = (time.perf_counter() - t0) / reps * 1e3
    for _ in range(3): device.msm_window_sums(bases, scal)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): device.msm_window_sums(bases, scal)
    e1.record(); torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / reps
    _lib.profile_enable(True)
    for k in range(3): _lib.profile_collect(k)
    for _ in range(reps): device.msm_window_sums(bases, scal)
    torch.cuda.synchronize()
    ph = [_lib.profile_collect(k)[0] / reps for k in range(3)]
    _lib.profile_enable(False)
    print(f"lg={lg} plan={device.msm_plan(n)} total {total:.3f} ms  device {dev_ms:.3f} ms  sort {ph[0]:.3f} acc {ph[1]:.3f} reduce {ph[2]:.3f}", flush=True)
    del bases, scal
