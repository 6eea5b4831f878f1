tag=${1:-ll}; prec=${2:-i8x5}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --precision $prec > gpurun_out/${tag}_ncu.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/${tag}_launches.csv')) if len(r)>5]
hdr=rows[0]; ci={h:i for i,h in enumerate(hdr)}
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[ci['Metric Value']].replace(',',''))
    except: continue
    k=r[ci['Kernel Name']][:50]; agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
unit=rows[1][ci['Metric Unit']]
for k,v in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:50s} n={v[0]:4d} total={v[1]:14.1f} {unit} {100*v[1]/tot:5.1f}%")
PY
