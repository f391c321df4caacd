# Round-2 GPU call AC: compute-sanitizer (memcheck, synccheck) over the quad-lane tail and the sonic commit pass at small sizes
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from snarkvm_b200 import device
for lg in (6, 9, 10, 12, 13):
    n = 1 << lg
    bases = device.generate_bases(n, 7)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    scal = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    scal[:, 3] &= (1 << 60) - 1
    a = device.msm(bases, scal)
    same = scal[:1].repeat(n, 1).contiguous()          # one hot bucket per window: folds
    b = device.msm(bases, same)
    polys = [scal[: n // 2].contiguous(), scal[: n // 3].contiguous()]
    c = device.sonic_commit_batch([bases, bases[5:]], polys)
    print(lg, a[:2], b[:2], c[0][:1], flush=True)
torch.cuda.synchronize()
print("done")
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san.py > gpurun_out/r2ac_memcheck.log 2>&1; echo memcheck rc=$?; tail -5 gpurun_out/r2ac_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 3 python /tmp/san.py > gpurun_out/r2ac_synccheck.log 2>&1; echo synccheck rc=$?; tail -5 gpurun_out/r2ac_synccheck.log
timeout 600 python -m pytest tests/test_sonic_gpu.py tests/test_varuna_gpu.py -m gpu -q -x > gpurun_out/r2ac_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2ac_pytest.log
timeout 900 python tools/bench_varuna.py prove 16 18 20 > gpurun_out/r2ac_varuna_prove.log 2>&1; echo rc=$?; cut -c1-700 gpurun_out/r2ac_varuna_prove.log
python tools/phase_sizes.py 10 12 14 16 > gpurun_out/r2ac_phases.log 2>&1; cat gpurun_out/r2ac_phases.log
