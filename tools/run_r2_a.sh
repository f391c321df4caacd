# Round-2 GPU call A: parity suite on the new MSM core, A/B of the pair kernels, first bench line, ncu of the new kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_ntt_gpu.py::test_ntt_large_sizes_vs_oracle -k "not full_size" > gpurun_out/r2a_pytest_fast.log 2>&1; echo rc=$?; tail -5 gpurun_out/r2a_pytest_fast.log
timeout 600 python tools/ab_v2.py > gpurun_out/r2a_ab_v2.log 2>&1; echo rc=$?; cat gpurun_out/r2a_ab_v2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo rc=$?; tail -c 3000 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_level2" -c 2 -f -o gpurun_out/r2a_pair2 python tools/time_sizes.py 24 > gpurun_out/r2a_ncu.log 2>&1; echo rc=$?
timeout 1200 python -m pytest tests -m gpu -q -k "full_size or large_sizes" > gpurun_out/r2a_pytest_full.log 2>&1; echo rc=$?; tail -5 gpurun_out/r2a_pytest_full.log
