# Round-2 GPU call V: per-kernel durations of a 2^12 MSM at two item caps (which kernel pays for extra item partials?)
mkdir -p gpurun_out
for cap in 8 64; do
SNARKVM_B200_MSM_CAP=$cap timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_launches_cap$cap.csv python tools/time_sizes.py 12 > /dev/null 2>&1
done
python - <<'PY'
import csv
for cap in (8, 64):
    rows = list(csv.reader(open(f"gpurun_out/r2v_launches_cap{cap}.csv")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = [(r[ki].split("(")[0].replace("void ", "").replace("b200::", ""), float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)) for r in data if len(r) > vi]
    starts = [i for i, (k, _) in enumerate(seq) if k.startswith("k_digits<0")]
    one = seq[starts[-2]:starts[-1]] if len(starts) >= 2 else seq
    print(f"--- cap={cap}: {len(one)} launches, {sum(t for _, t in one):.1f} us of kernel time")
    for k, t in one: print(f"   {t:8.1f} us  {k[:70]}")
PY
