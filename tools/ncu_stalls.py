"""Per-site stall summary of one kernel launch from an .ncu-rep source page (run in the build container):
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep <kernel regex> [launch index]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iex, ist = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
names = ['stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_math', 'stall_short_sb', 'stall_mio', 'stall_lg', 'stall_not_selected', 'stall_no_inst',
         'stall_dispatch', 'stall_branch_resolving', 'stall_selected']
cols = {k: hdr.index(k) for k in names}
def num(x):
    try: return int(x)
    except ValueError: return 0
seen, data = set(), []
for r in rows[2:]:
    if len(r) > max(cols.values()) and r[ia] not in seen and r[iex].isdigit():
        seen.add(r[ia]); data.append(r)
tot_s = sum(num(r[ist]) for r in data); tot_i = sum(num(r[iex]) for r in data)
print(rows[0][1][:100]); print('instructions', tot_i, 'samples', tot_s)
for k, i in cols.items():
    print(f"  {k:24s} {sum(num(r[i]) for r in data) / max(tot_s, 1) * 100:5.1f}%")
for key in ('stall_long_sb', 'stall_barrier', 'stall_short_sb', 'stall_mio'):
    print('--- top', key)
    for r in sorted(data, key=lambda r: -num(r[cols[key]]))[:14]:
        print(f"  [{data.index(r):5d}] {r[isrc].strip()[:78]:78s} exec {r[iex]:>9s}  {num(r[cols[key]])}")
