import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from snarkvm_b200 import device
from snarkvm_b200.cuda import NTTDirection, NTTType
import sys
for lg in ([int(a) for a in sys.argv[1:]] or (16, 20, 22, 24)):
    n = 1 << lg
    x = torch.from_numpy(np.random.default_rng(0).integers(0, 2**60, size=(n, 4), dtype=np.int64)).cuda()
    sc = torch.empty_like(x)
    for d, t in ((NTTDirection.Forward, NTTType.Standard), (NTTDirection.Inverse, NTTType.Coset)):
        for _ in range(3): device.ntt_(x, d, t, sc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): device.ntt_(x, d, t, sc)
        e1.record(); torch.cuda.synchronize()
        print(f"lg={lg} dir={int(d)} type={int(t)}: {e0.elapsed_time(e1)/10:.3f} ms  {n/(e0.elapsed_time(e1)/10)/1e6:.2f} Gel/s")
