"""Per-kernel shares of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...).
usage: python tools/launch_shares.py X.csv > X_shares.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*$", "", r[ki].replace("<unnamed>::", "").replace("void ", ""))
    name = name.replace("(bool)", "").replace("(int)", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
print("kernel,launches,total_us,avg_us,share_pct")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('"%s",%d,%.1f,%.1f,%.2f' % (k, n, t, t / n, 100 * t / tot))
