# Round-2 GPU call N: pipelined snarkvm_ntt (parity + timing with the switch on/off), whole GPU suite, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ntt_gpu.py -m gpu -q -x -k "host_ffi" > gpurun_out/r2n_pytest_ntt.log 2>&1; echo rc=$?; tail -5 gpurun_out/r2n_pytest_ntt.log
cat > /tmp/ntt_e2e.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from snarkvm_b200 import cuda as shim
for lg in (20, 22, 24):
    n = 1 << lg
    x = np.random.default_rng(1).integers(0, 2**60, size=(n, 4), dtype=np.uint64)
    pin_t = torch.from_numpy(x.view(np.int64).copy()).pin_memory(); pin = pin_t.numpy().view(np.uint64)
    for name, buf in (("pageable", x), ("pinned", pin)):
        shim.NTT(n, buf, shim.NTTInputOutputOrder.NN, shim.NTTDirection.Forward, shim.NTTType.Standard)
        t0 = time.perf_counter()
        for _ in range(5): shim.NTT(n, buf, shim.NTTInputOutputOrder.NN, shim.NTTDirection.Forward, shim.NTTType.Standard)
        print(f"lg={lg} {name} pipeline={'off' if os.environ.get('SNARKVM_B200_NTT_NO_PIPELINE') else 'on'} {(time.perf_counter()-t0)*200:.2f} ms", flush=True)
PY
python /tmp/ntt_e2e.py > gpurun_out/r2n_ntt_e2e.log 2>&1; SNARKVM_B200_NTT_NO_PIPELINE=1 python /tmp/ntt_e2e.py >> gpurun_out/r2n_ntt_e2e.log 2>&1; cat gpurun_out/r2n_ntt_e2e.log
timeout 1800 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2n_pytest.log 2>&1; echo rc=$?; tail -12 gpurun_out/r2n_pytest.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo rc=$?; tail -c 1000 gpurun_out/r2n_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print({k: d[k] for k in ('value','ms_per_step','checked')}, d['e2e']['ms_per_step'], d['e2e_pinned']['ms_per_step'], d['ntt']['ms_per_transform'], d['ntt']['e2e']['ms_per_step'], d['ntt']['e2e_pinned']['ms_per_step'], d['checks'])
PY
