"""8 × 2^lg-coefficient commitments over one resident SRS in one pass (sonic_pc/mod.rs:177-257), the plan's window size from the
environment: SNARKVM_B200_MSM_C / _LEVELS for A/B runs:   python tools/time_batch.py 20   (on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from snarkvm_b200 import device
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
powers = device.generate_bases(n, 7)
g = torch.Generator(device="cuda"); g.manual_seed(1)
polys = []
for i in range(8):
    c = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g); c[:, 3] &= (1 << 60) - 1
    polys.append(c)
for _ in range(3): device.kzg_commit_batch(powers, polys)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): device.kzg_commit_batch(powers, polys)
ms = (time.perf_counter() - t0) / 5 * 1e3
print(f"8 x 2^{lg} C={os.environ.get('SNARKVM_B200_MSM_C')} LEVELS={os.environ.get('SNARKVM_B200_MSM_LEVELS')}: {ms:.2f} ms  {8 * n / ms / 1e3:.1f} M coeff/s", flush=True)
