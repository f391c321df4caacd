# Round-2 GPU call U: item cap sweep on the quad path (reduce-phase time vs items per bucket), lockstep quad loops
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -q -x -k "host_ffi or unequal or edge or degenerate or skewed or real_srs or device_api or quad or half_repeated or batch_one_pass or window_sums or precomputed_bases_vs or kzg_commit_vs" > gpurun_out/r2u_pytest.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2u_pytest.log
for cap in 4 8 16 32; do echo "CAP=$cap"; SNARKVM_B200_MSM_CAP=$cap python tools/phase_sizes.py 12 13 14 15 16; done > gpurun_out/r2u_cap2.log 2>&1
cat gpurun_out/r2u_cap2.log
python tools/phase_sizes.py 6 8 9 10 11 > gpurun_out/r2u_small.log 2>&1; cat gpurun_out/r2u_small.log
