# Round-2 final GPU call: smoke, whole GPU suite, default bench
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"; echo smoke rc=$?
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2final_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2final_pytest.log
timeout 900 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; echo rc=$?; tail -c 300 gpurun_out/r2final_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2final_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'checked', 'gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], all(d['checks'].values()))
PY
