cd $GRAFT_REPO_ROOT
AZN_NMS_NO_PARTITION=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2al_nms2k.csv python tools/microbench.py --only nms --sizes 2000 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2al_nms2k.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]
out=[]
for r in rows[hi+1:]:
    d=dict(zip(h,r))
    out.append((d["Kernel Name"][:60], d["Metric Value"]))
for k,v in out[-14:]: print(k, v)
PY
