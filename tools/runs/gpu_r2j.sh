set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2j_launches_mma.csv python bench.py --steps 2 --warmup 3 --no-graph --no-extra --no-cpu-baseline --heads mma > gpurun_out/r2j_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2j_launches_gemm.csv python bench.py --steps 2 --warmup 3 --no-graph --no-extra --no-cpu-baseline --heads gemm > gpurun_out/r2j_b.log 2>&1
python - <<'PY'
import csv
for tag in ("mma","gemm"):
    rows=list(csv.reader(open("gpurun_out/r2j_launches_%s.csv"%tag)))
    hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
    h=rows[hi]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
    data=[(r[kn][:48], float(r[mv].replace(",",""))) for r in rows[hi+1:] if len(r)>mv]
    print(tag, len(data))
    for n,v in data[-36:]: print("  ", round(v/1000,1), n)
PY
