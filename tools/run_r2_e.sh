set -x
timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -q -k "pair_levels or half_repeated or precomputed_bases_vs or batch" > gpurun_out/r2e_pytest.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2e_pytest.log
SNARKVM_B200_MSM_PAIR_MINB=3 timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -q -k "pair_levels or half_repeated" > gpurun_out/r2e_pytest3.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2e_pytest3.log
timeout 600 python tools/ab_v2.py 24 22 > gpurun_out/r2e_ab.log 2>&1; echo rc=$?; cat gpurun_out/r2e_ab.log
SNARKVM_B200_MSM_PAIR_MINB=3 SNARKVM_B200_MSM_SCRATCH_GB=40 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_level2" -c 1 -f -o gpurun_out/r2e_pair3 python tools/time_sizes.py 24 > gpurun_out/r2e_ncu3.log 2>&1; echo rc=$?
