"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST 1/parts of the run."""
import csv, collections, re, sys
path = sys.argv[1]; parts = int(sys.argv[2]) if len(sys.argv) > 2 else 1; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
n = len(data); last = data[(parts - 1) * n // parts:]
agg = collections.OrderedDict()
for r in last:
    name = re.sub(r"\(.*", "", r[ki])[:100]
    t = float(r[vi].replace(",", ""))
    t = t / 1e6 if r[ui] == "ns" else (t / 1e3 if r[ui] == "us" else t)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print("launches %d, total %.3f ms" % (sum(a[0] for a in agg.values()), tot))
print("| ms | share | launches | kernel |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("| %.3f | %.1f %% | %d | `%s` |" % (t, 100 * t / tot, c, k))
