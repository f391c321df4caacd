# Round-2 GPU call P: smoke(), launch lists of small MSMs (where does a 2^12 / 2^16 MSM spend its time?)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"; echo smoke rc=$?
for lg in 12 16; do timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_$lg.csv python tools/time_sizes.py $lg > /dev/null 2>&1; done
python - <<'PY'
import csv, collections
for lg in (12, 16):
    rows = list(csv.reader(open(f"gpurun_out/r2p_launches_{lg}.csv")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = [(r[ki].split("(")[0].replace("void ", "").replace("b200::", ""), float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)) for r in data if len(r) > vi]
    # the last complete MSM = launches after the last k_generate_bases, split per call by the first k_digits<0,...>
    starts = [i for i, (k, _) in enumerate(seq) if k.startswith("k_digits<0")]
    one = seq[starts[-2]:starts[-1]] if len(starts) >= 2 else seq
    print(f"--- lg={lg}: {len(one)} launches, {sum(t for _, t in one):.1f} us of kernel time")
    for k, t in one: print(f"   {t:8.1f} us  {k[:70]}")
PY
