"""Ad-hoc large-size sanity run on the GPU: 2^26-point MSM (closed form against the oracle) and 2^26 / 2^27 NTT round trips."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import bls12_377 as py, cpu
from helpers import affine_array, generated_base_multipliers, random_canonical_fr
from snarkvm_b200 import device
from snarkvm_b200.cuda import NTTDirection, NTTType
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << lg
t = time.time(); bases = device.generate_bases(n, 4242); torch.cuda.synchronize(); print("generate", time.time() - t)
scal = random_canonical_fr(n, 11)
dscal = torch.from_numpy(scal.view(np.int64)).cuda()
print("plan", device.msm_plan(n))
t = time.time(); got = device.msm(bases, dscal); print("msm s", time.time() - t)
t = time.time(); got = device.msm(bases, dscal); print("msm s", time.time() - t)
ks = np.zeros((n, 4), dtype=np.uint64); ks[:, 0] = generated_base_multipliers(4242, n)
want = cpu.g1_mul(affine_array([py.G1_GENERATOR])[0], cpu.fr_dot_canonical(scal, ks))
print("MSM 2^%d closed form:" % lg, bool((got == want).all()))
del bases, dscal; torch.cuda.empty_cache()
for l2 in (lg, lg + 1):
    x = torch.from_numpy(random_canonical_fr(1 << l2, 5).view(np.int64)).cuda()
    y = device.ntt_(x.clone(), NTTDirection.Forward, NTTType.Coset)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); z = device.ntt_(y, NTTDirection.Inverse, NTTType.Coset); e1.record(); torch.cuda.synchronize()
    print("NTT 2^%d coset round trip:" % l2, bool(torch.equal(z, x)), "inverse ms", e0.elapsed_time(e1))
    del x, y, z; torch.cuda.empty_cache()
