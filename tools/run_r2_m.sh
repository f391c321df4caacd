# Round-2 GPU call M (2 GPUs): sharded MSM parity over NCCL, bench at N = 2 (weak + strong legs, e2e on host buffers)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_next_rows_gpu.py -m gpu -q -x -k "sharded or evaluations_type" > gpurun_out/r2m_pytest.log 2>&1; echo rc=$?; tail -6 gpurun_out/r2m_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo rc=$?; tail -c 1500 gpurun_out/r2m_bench_n2.err; cat gpurun_out/r2m_bench_n2.json
