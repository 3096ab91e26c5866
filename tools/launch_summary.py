"""per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv)"""
import csv, re, sys
from collections import OrderedDict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[kn])
    name = re.sub(r"^void ", "", name)[:90]
    t = float(r[mv].replace(",", "")) / 1e6
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] = t
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':92s} {'launches':>8s} {'total ms':>10s} {'last ms':>9s} {'share':>6s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:92s} {a[0]:8d} {a[1]:10.3f} {a[2]:9.3f} {100 * a[1] / tot:5.1f}%")
