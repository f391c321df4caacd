"""snarkvm_msm on a registered slice, plain vs precomputed tables: python tools/time_registered.py [lg]   (run on the GPU box)"""
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
torch.cuda.empty_cache()
for name, reg in (("registered", shim.register_bases), ("registered+tables", shim.register_bases_precomputed)):
    t0 = time.perf_counter(); reg(b_np); setup = time.perf_counter() - t0
    ok = bool((shim.msm(b_np, s_np) == ref).all())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): shim.msm(b_np, s_np)
    ms = (time.perf_counter() - t0) * 1e3 / 3
    shim.unregister_bases(b_np)
    print(f"lg={lg} {name}: ok={ok} e2e={ms:.1f} ms setup={setup:.2f} s", flush=True)
