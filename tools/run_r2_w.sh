# Round-2 GPU call W: SonicKZG10 on the device vs oracle/sonic.py, Varuna hiding mode + linear combinations + openings
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sonic_gpu.py tests/test_varuna_gpu.py -m gpu -q -x > gpurun_out/r2w_pytest_sonic.log 2>&1; echo rc=$?; tail -25 gpurun_out/r2w_pytest_sonic.log
