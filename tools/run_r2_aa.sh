# Round-2 GPU call AA: non-temporal staging copies — host-buffer parity, pageable e2e of snarkvm_msm / snarkvm_ntt at N = 1
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msm_gpu.py tests/test_ntt_gpu.py -m gpu -q -x -k "host_ffi or chunked or registered or concurrent or ffi" > gpurun_out/r2aa_pytest.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2aa_pytest.log
python tools/time_e2e_pageable.py 24 > gpurun_out/r2aa_e2e.log 2>&1; cat gpurun_out/r2aa_e2e.log
python /dev/stdin <<'PY' >> gpurun_out/r2aa_e2e.log 2>&1
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from snarkvm_b200 import cuda as shim
for lg in (22, 24):
    n = 1 << lg
    x = np.random.default_rng(1).integers(0, 2**60, size=(n, 4), dtype=np.uint64)
    shim.NTT(n, x, shim.NTTInputOutputOrder.NN, shim.NTTDirection.Forward, shim.NTTType.Standard)
    t0 = time.perf_counter()
    for _ in range(5): shim.NTT(n, x, shim.NTTInputOutputOrder.NN, shim.NTTDirection.Forward, shim.NTTType.Standard)
    print(f"snarkvm_ntt lg={lg} pageable {(time.perf_counter()-t0)*200:.2f} ms", flush=True)
PY
tail -3 gpurun_out/r2aa_e2e.log
