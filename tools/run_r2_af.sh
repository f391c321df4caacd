# Round-2 GPU call AF: per-kernel durations of the reduce phase of a 2^24-point MSM (final code)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2af_launches_24.csv python tools/time_sizes.py 24 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2af_launches_24.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = [(r[ki].split("(")[0].replace("void ", "").replace("b200::", ""), float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)) for r in data if len(r) > vi]
starts = [i for i, (k, _) in enumerate(seq) if k.startswith("k_digits<0")]
one = seq[starts[-2]:starts[-1]] if len(starts) >= 2 else seq
print(f"--- 2^24: {len(one)} launches, {sum(t for _, t in one) / 1e3:.2f} ms of kernel time")
for k, t in one: print(f"   {t:10.1f} us  {k[:80]}")
PY
