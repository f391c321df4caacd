# Round-2 GPU call J: whole suite; table MSM over scattered records (sweep); NTT old/new A/B; Varuna rounds bench
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2j_pytest.log 2>&1; echo rc=$?; tail -16 gpurun_out/r2j_pytest.log
SNARKVM_B200_LIB=$PWD/tools/bin/libsnarkvm_b200_nttold.so timeout 300 python tools/time_ntt.py > gpurun_out/r2j_ntt_old.log 2>&1; echo rc=$?; cat gpurun_out/r2j_ntt_old.log
timeout 300 python tools/time_ntt.py > gpurun_out/r2j_ntt_new.log 2>&1; echo rc=$?; cat gpurun_out/r2j_ntt_new.log
SNARKVM_B200_LIB=$PWD/tools/bin/libsnarkvm_b200_nttold.so timeout 300 python tools/time_ntt.py > gpurun_out/r2j_ntt_old2.log 2>&1; echo rc=$?; grep "lg=24" gpurun_out/r2j_ntt_old2.log
timeout 300 python tools/time_ntt.py > gpurun_out/r2j_ntt_new2.log 2>&1; echo rc=$?; grep "lg=24" gpurun_out/r2j_ntt_new2.log
timeout 900 python tools/tune_precomputed.py 24 20,22 3,4,5 > gpurun_out/r2j_tune_pre.log 2>&1; echo rc=$?; cat gpurun_out/r2j_tune_pre.log
timeout 600 python tools/tune_precomputed.py 22 19,20 3,4 >> gpurun_out/r2j_tune_pre.log 2>&1; timeout 600 python tools/tune_precomputed.py 20 16,17 2,3 >> gpurun_out/r2j_tune_pre.log 2>&1; tail -8 gpurun_out/r2j_tune_pre.log
timeout 600 python tools/ab_v2.py 20 21 > gpurun_out/r2j_ab.log 2>&1; echo rc=$?; cat gpurun_out/r2j_ab.log
timeout 900 python tools/bench_varuna.py 16 18 20 > gpurun_out/r2j_varuna.log 2>&1; echo rc=$?; cat gpurun_out/r2j_varuna.log
