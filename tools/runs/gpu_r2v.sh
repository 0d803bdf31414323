set -x
cd $GRAFT_REPO_ROOT
for mode in 22 30; do
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none --csv --log-file gpurun_out/r2v_l$mode.csv python tools/microbench.py --only roi_pool --pool-mode $mode --sizes 20000 --hw 38,63 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r2v_l$mode.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]
seen={}
for r in rows[hi+1:]:
    d=dict(zip(h,r))
    k=d["Kernel Name"][:50]
    if "roi_" not in k: continue
    seen.setdefault(k,{}).setdefault(d["Metric Name"],[]).append(d["Metric Value"])
for k,v in seen.items():
    print(k)
    for m,vals in v.items(): print("   ",m,vals[-6:])
PY
done
