# Round-2 GPU call H: Karatsuba multiplier (microbench vs schoolbook, self-test), chunked record scatter, windowed G1 iFFT
set -x
mkdir -p gpurun_out
tools/bin/ff_microbench_s > gpurun_out/r2h_ffbench_schoolbook.log 2>&1; echo rc=$?; grep -E "tpb=128 blocks/SM=4|tpb=512 blocks/SM=4|checks|PASS|FAIL" gpurun_out/r2h_ffbench_schoolbook.log
tools/bin/ff_microbench_k > gpurun_out/r2h_ffbench_karatsuba.log 2>&1; echo rc=$?; grep -E "tpb=128 blocks/SM=4|tpb=512 blocks/SM=4|checks|PASS|FAIL" gpurun_out/r2h_ffbench_karatsuba.log
timeout 900 python -m pytest tests/test_msm_gpu.py tests/test_next_rows_gpu.py tests/test_ntt_gpu.py -m gpu -q -x -k "not full_size and not large_sizes and not concurrent_large" > gpurun_out/r2h_pytest.log 2>&1; echo rc=$?; tail -8 gpurun_out/r2h_pytest.log
timeout 600 python tools/time_g1_ntt.py 12 14 16 18 > gpurun_out/r2h_g1ntt.log 2>&1; echo rc=$?; cat gpurun_out/r2h_g1ntt.log
timeout 600 python tools/ab_v2.py 24 22 21 20 19 18 > gpurun_out/r2h_ab.log 2>&1; echo rc=$?; cat gpurun_out/r2h_ab.log
timeout 300 python tools/time_ntt.py > gpurun_out/r2h_ntt.log 2>&1; echo rc=$?; tail -12 gpurun_out/r2h_ntt.log
