# Round-2 GPU call C: lazy descriptors + warp-cooperative inversion
set -x
timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -q -k "selftest or pair_levels or half_repeated or precomputed_bases_vs or batch" > gpurun_out/r2c_pytest.log 2>&1; echo rc=$?; tail -8 gpurun_out/r2c_pytest.log
timeout 600 python tools/ab_v2.py 24 22 20 > gpurun_out/r2c_ab.log 2>&1; echo rc=$?; cat gpurun_out/r2c_ab.log
SNARKVM_B200_MSM_SCRATCH_GB=40 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_level2" -c 2 -f -o gpurun_out/r2c_pair python tools/time_sizes.py 24 > gpurun_out/r2c_ncu.log 2>&1; echo rc=$?
