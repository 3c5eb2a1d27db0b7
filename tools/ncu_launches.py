"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...) per kernel:
   python tools/ncu_launches.py gpurun_out/X.csv [last_n_steps] > profiles/X.launches.txt
Per-launch times are cold-cache and serialised (kernels that overlap on two streams in a real step are listed one after
the other): compare SHARES, not absolutes."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
per = collections.OrderedDict()
for r in rows[start + 1:]:
    rec = dict(zip(hdr, r))
    if "Metric Value" not in rec:
        continue
    v = float(rec["Metric Value"].replace(",", ""))
    u = rec["Metric Unit"]
    v = v / 1e3 if u == "us" else v / 1e6 if u == "ns" else v * 1e3 if u == "s" else v
    per.setdefault(rec["Kernel Name"].split("(")[0], []).append((v, rec["Grid Size"], rec["Block Size"]))
tot = sum(v for k in per for v, _, _ in per[k])
print("%-42s %5s %12s %10s %8s  %s" % ("kernel", "n", "total ms", "last ms", "share", "grid x block (last)"))
for k, L in per.items():
    print("%-42s %5d %12.3f %10.3f %7.1f%%  %s x %s" % (k[:42], len(L), sum(v for v, _, _ in L), L[-1][0], 100 * sum(v for v, _, _ in L) / tot, L[-1][1], L[-1][2]))
# steady-state step = the launches after the last far_kernel-containing assembly began
names = [(k, v) for k in per for v, _, _ in per[k]]
