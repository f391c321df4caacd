# Round-2 GPU call G: record-scatter sort — parity, A/B, ncu of every MSM kernel at 2^24
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -q -x -k "record_scatter or pair_levels or half_repeated or batch or hiding or device_api or full_size" > gpurun_out/r2g_pytest.log 2>&1; echo rc=$?; tail -8 gpurun_out/r2g_pytest.log
timeout 600 python tools/ab_v2.py 24 22 21 20 > gpurun_out/r2g_ab.log 2>&1; echo rc=$?; cat gpurun_out/r2g_ab.log
SNARKVM_B200_MSM_SCRATCH_GB=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_level2|k_scatter_records|k_bucket_accumulate|k_digits" -c 12 -f -o gpurun_out/r2g_msm python tools/time_sizes.py 24 > gpurun_out/r2g_ncu.log 2>&1; echo rc=$?
