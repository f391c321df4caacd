set -x
timeout 300 python bench.py > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --skip-cpu --e2e-steps 1 > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch.err; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bucket_accumulate -s 1 -c 1 -f -o gpurun_out/prof_msm_acc_r1 python bench.py --steps 1 --skip-cpu --skip-ntt --e2e-steps 1 > /dev/null 2> gpurun_out/ncu_acc.err; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 3 -c 3 -f -o gpurun_out/prof_ntt_pass_r1 python bench.py --lg 20 --steps 1 --skip-cpu --e2e-steps 1 > /dev/null 2> gpurun_out/ncu_ntt.err; echo rc=$?
cat gpurun_out/bench_r1a.json
ls -la gpurun_out
