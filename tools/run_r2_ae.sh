# Round-2 GPU call AE: quad-lane combine levels behind the per-chunk bucket reduction of large bucket sets — parity, then A/B timing
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_msm_gpu.py -m gpu -q -x -k "full_size or pair_levels or record_scatter or half_repeated or batch_large or kzg_commit_hiding_full or precomputed_bases_full or device_api" > gpurun_out/r2ae_pytest.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2ae_pytest.log
for q in 1 0; do echo "QUAD_TAIL=$q"; SNARKVM_B200_MSM_QUAD_TAIL=$q python tools/phase_sizes.py 19 20 22 24; done > gpurun_out/r2ae_phases.log 2>&1; cat gpurun_out/r2ae_phases.log
