# Round-2 GPU call O: NTT twiddle prefetch A/B + parity
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ntt_gpu.py -m gpu -q -x -k "not full_size and not large_sizes" > gpurun_out/r2o_pytest.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2o_pytest.log
for i in 1 2; do SNARKVM_B200_LIB=$PWD/tools/bin/libsnarkvm_b200_nttnopf.so timeout 300 python tools/time_ntt.py 20 22 24 2>&1 | sed "s/^/noprefetch  /"; timeout 300 python tools/time_ntt.py 20 22 24 2>&1 | sed "s/^/prefetch    /"; done > gpurun_out/r2o_ntt_prefetch.log; cat gpurun_out/r2o_ntt_prefetch.log
timeout 600 python -m pytest tests/test_ntt_gpu.py -m gpu -q -x -k "large_sizes" > gpurun_out/r2o_pytest2.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2o_pytest2.log
