import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
h=next(i for i,r in enumerate(rows) if "Kernel Name" in r)
hd=rows[h]; ki=hd.index("Kernel Name"); mi=hd.index("Metric Name"); vi=hd.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[h+1:]:
    if len(r)>vi and r[mi]=="gpu__time_duration.sum": d[r[ki].split("(")[0][-32:]].append(float(r[vi])/1e6)
for k,v in d.items(): print(sys.argv[1][-12:], k, "n=%d"%len(v), "sum_ms=%.2f"%sum(v), "first4=",[round(x,2) for x in v[:4]])
