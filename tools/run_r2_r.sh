# Round-2 GPU call R: warm per-phase times of small / mid MSMs
set -x
mkdir -p gpurun_out
python tools/phase_sizes.py 10 12 14 16 17 18 19 20 21 > gpurun_out/r2r_phases.log 2>&1; cat gpurun_out/r2r_phases.log
