# Round-2 GPU call Y: whole GPU suite, smoke, default bench on the current HEAD
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"; echo smoke rc=$?
timeout 1800 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2y_pytest.log 2>&1; echo rc=$?; tail -12 gpurun_out/r2y_pytest.log
timeout 1500 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo rc=$?; tail -c 800 gpurun_out/r2y_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2y_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'checked', 'gpu_launches')})
print('e2e', d['e2e']['ms_per_step'], 'pinned', d['e2e_pinned']['ms_per_step'])
print('sizes', {k: round(v['ms_per_msm'], 3) for k, v in (d.get('msm_sizes') or {}).items()})
print('ntt', d['ntt']['ms_per_transform'], d['ntt']['e2e']['ms_per_step'])
print('varuna', d['varuna']['s_per_proof'], d['varuna'].get('prove', {}).get('s_per_proof'))
print('checks', d['checks'])
PY
