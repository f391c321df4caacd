# Final round-1 measurement script (run under gpurun from the repo root).  Outputs land in gpurun_out/.
set -x
timeout 500 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo rc=$?
# every launch of the same command with its device time (cold-cache, serialised: compare shares)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-kzg --e2e-steps 1 > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch.err; echo rc=$?
# the dominant kernels once each with the full set (the accumulation phase of one 2^24-point MSM)
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_pair_level|k_bucket_accumulate" -s 5 -c 5 -f -o gpurun_out/prof_msm_r1 python bench.py --steps 1 --skip-cpu --skip-ntt --skip-kzg --e2e-steps 1 > /dev/null 2> gpurun_out/ncu_msm.err; echo rc=$?
python -c "import __graft_entry__ as g; g.smoke()"; echo smoke rc=$?
cat gpurun_out/bench_r1.json
