# Round-2 GPU call S: quad latency path — parity on small sizes, then warm phase times with the path on and off
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -q -x -k "host_ffi or unequal or edge or degenerate or skewed or real_srs or device_api or quad or half_repeated or batch_one_pass or window_sums" > gpurun_out/r2s_pytest.log 2>&1; echo rc=$?; tail -8 gpurun_out/r2s_pytest.log
python tools/phase_sizes.py 8 10 11 12 13 14 16 18 > gpurun_out/r2s_phases_quad.log 2>&1; cat gpurun_out/r2s_phases_quad.log
SNARKVM_B200_MSM_QUAD=0 python tools/phase_sizes.py 8 10 11 12 13 14 16 18 > gpurun_out/r2s_phases_noquad.log 2>&1; cat gpurun_out/r2s_phases_noquad.log
