# Round-2 GPU call K: G2 MSM parity, sparse mat-vec with hot rows, NTT variant A/B, Varuna bench, bench line
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_msm_g2_gpu.py tests/test_next_rows_gpu.py tests/test_varuna_gpu.py tests/test_msm_gpu.py -m gpu -q -x --durations=6 -k "not full_size" > gpurun_out/r2k_pytest.log 2>&1; echo rc=$?; tail -14 gpurun_out/r2k_pytest.log
for v in nttold nttswz ntttw; do SNARKVM_B200_LIB=$PWD/tools/bin/libsnarkvm_b200_$v.so timeout 300 python tools/time_ntt.py 2>&1 | grep "lg=24\|lg=22 dir=0" | sed "s/^/$v  /"; done > gpurun_out/r2k_ntt_variants.log; timeout 300 python tools/time_ntt.py 2>&1 | grep "lg=24\|lg=22 dir=0" | sed "s/^/both  /" >> gpurun_out/r2k_ntt_variants.log; cat gpurun_out/r2k_ntt_variants.log
timeout 900 python tools/bench_varuna.py 16 18 20 > gpurun_out/r2k_varuna.log 2>&1; echo rc=$?; cat gpurun_out/r2k_varuna.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo rc=$?; tail -c 1500 gpurun_out/r2k_bench.err; cat gpurun_out/r2k_bench.json
