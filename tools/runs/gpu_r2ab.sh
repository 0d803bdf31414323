set -x
cd $GRAFT_REPO_ROOT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ab_nms.csv python tools/microbench.py --only nms --sizes 20000 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2ab_nms.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]
seen={}
for r in rows[hi+1:]:
    d=dict(zip(h,r))
    k=d["Kernel Name"][:60]
    seen.setdefault(k,[]).append(float(d["Metric Value"]))
for k,v in seen.items():
    print(k, len(v), "last:", v[-3:], d["Metric Unit"])
PY
