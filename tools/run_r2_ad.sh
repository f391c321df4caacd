# Round-2 GPU call AB (4 GPUs): bench at N = 4 with the non-temporal staging copies (weak + strong legs, pageable e2e)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2ad_bench_n4.json 2> gpurun_out/r2ad_bench_n4.err; echo rc=$?; tail -c 600 gpurun_out/r2ad_bench_n4.err
python - <<'PY'
import json
for line in open('gpurun_out/r2ad_bench_n4.json'):
    if line.startswith('{'):
        d = json.loads(line)
        print({k: d[k] for k in ('value', 'ms_per_step', 'checked')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'pinned', d['e2e_pinned']['ms_per_step'], 'strong', d['sharded_total']['ms_per_step'], 'ntt e2e', d['ntt']['e2e']['ms_per_step'])
PY
