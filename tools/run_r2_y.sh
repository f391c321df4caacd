# Round-2 GPU call Y: whole GPU suite, smoke, default bench on the current HEAD
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"; echo smoke rc=$?
timeout 1800 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2y_pytest.log 2>&1; echo rc=$?; tail -12 gpurun_out/r2y_pytest.log
timeout 1500 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo rc=$?; tail -c 800 gpurun_out/r2y_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2y_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'checked', 'gpu_launches')})
print('e2e', d['e2e']['ms_per_step'], 'pinned', d['e2e_pinned']['ms_per_step'])
print('sizes', {k: round(v['ms_per_msm'], 3) for k, v in (d.get('msm_sizes') or {}).items()})
print('ntt', d['ntt']['ms_per_transform'], d['ntt']['e2e']['ms_per_step'])
print('varuna', d['varuna']['s_per_proof'], d['varuna'].get('prove', {}).get('s_per_proof'))
print('checks', d['checks'])
PY
python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from snarkvm_b200 import device
rng = np.random.default_rng(1)
from oracle import cpu, bls12_377 as py
# host tail alone: Horner over 32 window sums of c = 8 (what a 2^12-point call does after its D2H)
g = np.frombuffer(py.affine_bytes(py.G1_GENERATOR), dtype=np.uint8)
pts = []
for k in range(32):
    proj = cpu.g1_mul(g, np.array([k + 3, 0, 0, 0], dtype=np.uint64))
    x, y = proj[:6], proj[6:12]
    one = proj[12:18]
    pts.append(np.concatenate([x, y, one, one]))          # XYZZ with ZZ = ZZZ = 1
sums = np.stack(pts).astype(np.uint64)
device.msm_finish(sums, 8)
t0 = time.perf_counter()
for _ in range(200): device.msm_finish(sums, 8)
print(f"host Horner (32 windows, c = 8) incl. ctypes: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
PY
