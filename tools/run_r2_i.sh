# Round-2 GPU call I: whole GPU suite on the new build, latency path A/B, NTT, e2e pageable sweep, G1 iFFT, table precompute
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2i_pytest.log 2>&1; echo rc=$?; tail -16 gpurun_out/r2i_pytest.log
timeout 600 python tools/ab_v2.py 24 22 20 19 > gpurun_out/r2i_ab.log 2>&1; echo rc=$?; cat gpurun_out/r2i_ab.log
timeout 300 python tools/time_ntt.py > gpurun_out/r2i_ntt.log 2>&1; echo rc=$?; cat gpurun_out/r2i_ntt.log
timeout 600 python tools/time_g1_ntt.py 12 16 20 > gpurun_out/r2i_g1ntt.log 2>&1; echo rc=$?; cat gpurun_out/r2i_g1ntt.log
timeout 900 python tools/time_e2e_pageable.py 24 > gpurun_out/r2i_e2e_pageable.log 2>&1; echo rc=$?; cat gpurun_out/r2i_e2e_pageable.log
timeout 600 python tools/time_precomputed.py > gpurun_out/r2i_precomputed.log 2>&1; echo rc=$?; tail -12 gpurun_out/r2i_precomputed.log
