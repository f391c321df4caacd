# Round-2 GPU call L: divide_by_vanishing / sparse mat-vec parity, NTT variant A/B, Varuna 2^20, launch list + ncu --set full of the final kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_next_rows_gpu.py tests/test_varuna_gpu.py -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo rc=$?; tail -5 gpurun_out/r2l_pytest.log
for v in nttold nttswz ntttw; do SNARKVM_B200_LIB=$PWD/tools/bin/libsnarkvm_b200_$v.so timeout 300 python tools/time_ntt.py 2>&1 | grep "lg=24\|lg=22 dir=0\|rror" | sed "s/^/$v  /"; done > gpurun_out/r2l_ntt_variants.log; timeout 300 python tools/time_ntt.py 2>&1 | grep "lg=24\|lg=22 dir=0" | sed "s/^/both  /" >> gpurun_out/r2l_ntt_variants.log; cat gpurun_out/r2l_ntt_variants.log
timeout 600 python tools/bench_varuna.py 20 > gpurun_out/r2l_varuna.log 2>&1; echo rc=$?; cat gpurun_out/r2l_varuna.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-kzg --e2e-steps 1 --varuna-lg 0 --g2-lg 0 --skip-strong --skip-registered > gpurun_out/r2l_ncu_launch_bench.json 2> gpurun_out/r2l_ncu_launch.err; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pair_level2|k_bucket_accumulate_dense|k_scatter_records|k_digits" -c 9 -f -o gpurun_out/r2l_msm python tools/time_sizes.py 24 > gpurun_out/r2l_ncu_msm.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ntt_pass" -s 3 -c 3 -f -o gpurun_out/r2l_ntt python tools/time_ntt.py 24 > gpurun_out/r2l_ncu_ntt.log 2>&1; echo rc=$?
