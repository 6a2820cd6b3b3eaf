mkdir -p gpurun_out
MMF_BENCH_ALLOW_SHORT=1 timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_c2.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/launches_bench_c2.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; ci={n:i for i,n in enumerate(h)}
agg=collections.defaultdict(lambda:[0,0.0]); seq=[]
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    v=float(r[ci["Metric Value"]].replace(",","")); u=r[ci["Metric Unit"]]
    v*={"ns":1e-6,"us":1e-3,"ms":1.0}.get(u,1e-6)
    k=r[ci["Kernel Name"]][:64]; agg[k][0]+=1; agg[k][1]+=v; seq.append((k,v))
tot=sum(v[1] for v in agg.values()); print("total ms",tot,"launches",len(seq))
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:22]: print(f"  {v[1]:9.3f} ms x{v[0]:4d}  {k}")
PY
