import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    v = float(row['Metric Value']); u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    name = row['Kernel Name']
    name = name.replace('void cola::', '').replace('cola::', '')
    agg[name[:70]].append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} n={len(v):5d} mean={sum(v)/len(v):10.1f}us total={sum(v)/1e3:9.2f}ms share={sum(v)/tot*100:5.1f}%")
print(f"total {tot/1e3:.2f} ms")
