# Round-2 GPU call Z: launch lists of small MSMs with the quad-lane tail (final code) + full-metric capture of its kernels
set -x
mkdir -p gpurun_out
for lg in 12 16; do timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_$lg.csv python tools/time_sizes.py $lg > /dev/null 2>&1; done
python - <<'PY'
import csv
for lg in (12, 16):
    rows = list(csv.reader(open(f"gpurun_out/r2z_launches_{lg}.csv")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = [(r[ki].split("(")[0].replace("void ", "").replace("b200::", ""), float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)) for r in data if len(r) > vi]
    starts = [i for i, (k, _) in enumerate(seq) if k.startswith("k_digits<0")]
    one = seq[starts[-2]:starts[-1]] if len(starts) >= 2 else seq
    print(f"--- lg={lg}: {len(one)} launches, {sum(t for _, t in one):.1f} us of kernel time (cold, serialised)")
    for k, t in one: print(f"   {t:8.1f} us  {k[:70]}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bucket_reduce_quad|k_window_combine_quad|k_fold_hot_quad|k_combine_level_quad" -c 8 -o gpurun_out/r2z_quad python tools/time_sizes.py 16 > /dev/null 2>&1; ls -la gpurun_out/r2z_quad.ncu-rep
