# Round-2 GPU call X: prove-shaped Varuna timing (rounds + SonicKZG10 commits with bounds + linear combinations + openings)
set -x
mkdir -p gpurun_out
timeout 900 python tools/bench_varuna.py prove 16 18 20 > gpurun_out/r2x_varuna_prove.log 2>&1; echo rc=$?; cat gpurun_out/r2x_varuna_prove.log | cut -c1-1200
timeout 600 python tools/bench_varuna.py 16 18 20 > gpurun_out/r2x_varuna_rounds.log 2>&1; echo rc=$?; cat gpurun_out/r2x_varuna_rounds.log | cut -c1-900
