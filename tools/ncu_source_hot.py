"""Aggregate the ncu source page (ncu -i X.ncu-rep --page source --csv --print-source=cuda,sass > f.csv) by CUDA line.
usage: python tools/ncu_source_hot.py f.csv [topN]"""
import csv
import collections
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        fpath = rows[i][1]
        func = rows[i + 1][1][:80]
        hdr = rows[i + 2]
        j = i + 3
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j]); j += 1
        ix = {h: k for k, h in enumerate(hdr)}
        # first block is CUDA-source view: columns Line No, Source, ... ; sass rows have Address
        li, si, ai = hdr.index("Line No"), hdr.index("Source"), hdr.index("Address")
        ie = ix["Instructions Executed"]; ss = ix["# Samples"]
        agg = collections.defaultdict(lambda: [0, 0, ""])
        tot_i = tot_s = 0
        for r in body:
            if len(r) <= ie or not r[li]:
                continue
            try:
                n = int(float(r[ie] or 0)); smp = int(float(r[ss] or 0))
            except ValueError:
                continue
            a = agg[r[li]]
            a[0] += n; a[1] += smp; a[2] = r[si][:110]
            tot_i += n; tot_s += smp
        print(f"=== {func}\n    {fpath}  total inst {tot_i:.3e}  samples {tot_s}")
        for line, (n, smp, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"  L{line:>4s} inst {100 * n / max(tot_i, 1):5.1f}%  stall-samples {100 * smp / max(tot_s, 1):5.1f}%  {src.strip()}")
        i = j
    else:
        i += 1
