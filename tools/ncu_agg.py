import csv,re,collections,sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict()
for r in rows:
    v=float(r['Metric Value'].replace(',',''))/1000
    key=re.sub(r"\(.*","",r['Kernel Name'])[:60]+" grid="+r['Grid Size']
    agg.setdefault(key,[]).append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k:80s} n={len(v):4d} mean_us={sum(v)/len(v):8.2f} min={min(v):7.2f} max={max(v):7.2f} total_ms={sum(v)/1000:7.3f} share={sum(v)/tot:5.3f}")
print("total ms", tot/1000)
