# Round-2 GPU call AB (2 GPUs): bench at N = 2 with the non-temporal staging copies (weak + strong legs, pageable e2e)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 300 python -m pytest tests/test_sharded_gpu.py -m gpu -q -x > gpurun_out/r2ab_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2ab_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2ab_bench_n2.json 2> gpurun_out/r2ab_bench_n2.err; echo rc=$?; tail -c 600 gpurun_out/r2ab_bench_n2.err
python - <<'PY'
import json
for line in open('gpurun_out/r2ab_bench_n2.json'):
    if line.startswith('{'):
        d = json.loads(line)
        print({k: d[k] for k in ('value', 'ms_per_step', 'checked')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'pinned', d['e2e_pinned']['ms_per_step'], 'strong', d['sharded_total']['ms_per_step'], 'ntt e2e', d['ntt']['e2e']['ms_per_step'])
PY
