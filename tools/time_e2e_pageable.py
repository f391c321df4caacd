"""snarkvm_msm / snarkvm_ntt end to end with PAGEABLE host buffers for several copy-thread counts and upload ranges
(one subprocess per setting: the copy pool is created once per process):  python tools/time_e2e_pageable.py [lg]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import numpy as np, torch
from snarkvm_b200 import cuda as shim, device
lg = int(sys.argv[1]); n = 1 << lg
bases = device.generate_bases(n, 7)
rng = np.random.default_rng(0)
s = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); s[:, 3] &= np.uint64((1 << 60) - 1)
b_np = np.empty((n, 104), dtype=np.uint8); b_np[:] = bases.cpu().numpy()
s_np = s.copy()
ref = device.msm(bases, torch.from_numpy(s.view(np.int64)).cuda())
del bases
ok = bool((shim.msm(b_np, s_np) == ref).all())
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): shim.msm(b_np, s_np)
ms = (time.perf_counter() - t0) * 1e3 / 3
print(f"lg={lg} threads={os.environ.get('SNARKVM_B200_COPY_THREADS')} ranges={os.environ.get('SNARKVM_B200_MSM_CHUNKS')} ok={ok} msm e2e pageable {ms:.1f} ms", flush=True)
''' % ROOT
lg = sys.argv[1] if len(sys.argv) > 1 else "24"
for threads in ("6", "12", "24", "48"):
    for ranges in ("1:3:4", "1:3:12", "1:2:4:9"):
        env = dict(os.environ, SNARKVM_B200_COPY_THREADS=threads, SNARKVM_B200_MSM_CHUNKS=ranges)
        subprocess.run([sys.executable, "-c", CHILD, lg], env=env, check=False)
