"""lagrange_basis (iFFT over G1 points) timing at the given log2 sizes: python tools/time_g1_ntt.py 12 14 16   (run on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from snarkvm_b200 import device

for lg in [int(a) for a in sys.argv[1:]]:
    n = 1 << lg
    bases = device.generate_bases(n, 11)
    out = device.g1_ntt(bases, True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if lg <= 16 else 1
    e0.record()
    for _ in range(reps): out = device.g1_ntt(bases, True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    back = device.g1_ntt(out, False)
    ok = bool((back == bases).all())
    print(f"lg={lg} g1 ifft {ms:.2f} ms  ({n * max(lg, 1) / 2 / ms / 1e3:.2f} M butterflies/s)  fft(ifft(x)) == x: {ok}", flush=True)
