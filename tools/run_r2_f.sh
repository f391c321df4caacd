# Round-2 GPU call F (re-entry): whole GPU suite, pair-kernel A/B, bench line, launch list, ncu of the pair kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 600 python tools/ab_v2.py 24 22 20 > gpurun_out/r2f_ab.log 2>&1; echo rc=$?; cat gpurun_out/r2f_ab.log
timeout 1800 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2f_pytest.log 2>&1; echo rc=$?; tail -30 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo rc=$?; tail -c 2000 gpurun_out/r2f_bench.err; cat gpurun_out/r2f_bench.json
SNARKVM_B200_MSM_SCRATCH_GB=40 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_level2" -c 4 -f -o gpurun_out/r2f_pair python tools/time_sizes.py 24 > gpurun_out/r2f_ncu.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv python tools/time_sizes.py 24 > gpurun_out/r2f_launch.log 2>&1; echo rc=$?
