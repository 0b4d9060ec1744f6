import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[ii], r[ki][:60]), {})[r[mi]] = float(r[vi].replace(",", ""))
for (i, k), m in per.items():
    if "gemm" not in k.lower() and "cutlass" not in k.lower() and "nvjet" not in k.lower() and "xmma" not in k.lower(): continue
    print(f"#{i} {k}")
    for name, v in m.items(): print(f"      {name:60s} {v:18,.0f}")
