# Round-2 GPU call Q (re-entry): whole GPU suite + default bench on HEAD, then the small-MSM launch lists of call P
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2q_pytest.log 2>&1; echo rc=$?; tail -14 gpurun_out/r2q_pytest.log
timeout 1200 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo rc=$?; tail -c 600 gpurun_out/r2q_bench.err; cut -c1-1500 gpurun_out/r2q_bench.json
bash tools/run_r2_p.sh > gpurun_out/r2p.log 2>&1; tail -120 gpurun_out/r2p.log
