"""snarkvm_msm (host buffers, pinned) with 1…4 upload ranges: python tools/time_e2e.py [lg]   (run on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from snarkvm_b200 import cuda as shim, device

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << lg
bases = device.generate_bases(n, 7)
rng = np.random.default_rng(0)
s = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); s[:, 3] &= np.uint64((1 << 60) - 1)
bh = bases.cpu().pin_memory(); sh = torch.from_numpy(s.view(np.int64)).pin_memory()
b_np, s_np = bh.numpy(), sh.numpy().view(np.uint64)
ref = device.msm(bases, sh.cuda())
del bases
for k in sys.argv[2:] or ("1", "2", "1:3:4", "1:2:5", "1:3:5:7", "2:5:9", "1:3:4"):
    os.environ["SNARKVM_B200_MSM_CHUNKS"] = str(k)
    ok = bool((shim.msm(b_np, s_np) == ref).all())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): shim.msm(b_np, s_np)
    ms = (time.perf_counter() - t0) * 1e3 / 3
    print(f"lg={lg} ranges={k} ok={ok} e2e={ms:.1f} ms", flush=True)
