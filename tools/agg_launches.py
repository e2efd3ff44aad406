"""aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name"""
import csv, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines=[l for l in open(path) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0,0.0,[]])
for row in r:
    try: v=float(row['Metric Value'].replace(',',''))
    except Exception: continue
    unit=row['Metric Unit']
    v = v/1e3 if unit=='ns' else v*1e3 if unit=='ms' else v
    name=row['Kernel Name'].split('(')[0][:60]
    a=agg[name]; a[0]+=1; a[1]+=v; a[2]=max(a[2],v); a[3].append(round(v))
tot=sum(a[1] for a in agg.values())
print("| kernel | launches | total us | avg us | max us | share |\n|---|---|---|---|---|---|")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1]/a[0]:.1f} | {a[2]:.1f} | {a[1]/tot*100:.1f}% |")
if len(sys.argv) > 2:
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:4]: print(k, a[3][:24])
