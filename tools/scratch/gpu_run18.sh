mkdir -p gpurun_out
for M in 1048576 65536; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_big --csv --log-file gpurun_out/launches_big_$M.csv python tools/profile_c5.py $M > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open("gpurun_out/launches_big_$M.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; ci={n:i for i,n in enumerate(h)}
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    v=float(r[ci["Metric Value"]].replace(",","")); u=r[ci["Metric Unit"]]
    v*={"ns":1e-6,"us":1e-3,"ms":1.0}.get(u,1e-6)
    agg[r[ci["Kernel Name"]][:60]][0]+=1; agg[r[ci["Kernel Name"]][:60]][1]+=v
print("M=$M")
for k,v in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"  {v[1]/v[0]:9.3f} ms avg x{v[0]:3d}  {k}")
PY
done
