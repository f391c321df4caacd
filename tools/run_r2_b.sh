# Round-2 GPU call B: pair kernel with descriptors + staggered inversions
set -x
timeout 900 python -m pytest tests/test_msm_gpu.py tests/test_next_rows_gpu.py tests/test_varuna_gpu.py -m gpu -q -k "not full_size and not concurrent_large" > gpurun_out/r2b_pytest.log 2>&1; echo rc=$?; tail -15 gpurun_out/r2b_pytest.log
timeout 600 python tools/ab_v2.py > gpurun_out/r2b_ab.log 2>&1; echo rc=$?; cat gpurun_out/r2b_ab.log
SNARKVM_B200_MSM_SCRATCH_GB=40 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_level2|k_pair_desc" -c 4 -f -o gpurun_out/r2b_pair python tools/time_sizes.py 24 > gpurun_out/r2b_ncu.log 2>&1; echo rc=$?
