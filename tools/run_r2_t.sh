# Round-2 GPU call T: quad tail + scan-free folds — parity on small sizes, warm phase times
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -q -x -k "host_ffi or unequal or edge or degenerate or skewed or real_srs or device_api or quad or half_repeated or batch_one_pass or window_sums or precomputed_bases_vs or kzg_commit_vs" > gpurun_out/r2t_pytest.log 2>&1; echo rc=$?; tail -8 gpurun_out/r2t_pytest.log
python tools/phase_sizes.py 8 10 11 12 13 14 15 16 17 18 > gpurun_out/r2t_phases_quad.log 2>&1; cat gpurun_out/r2t_phases_quad.log
