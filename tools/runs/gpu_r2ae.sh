cd $GRAFT_REPO_ROOT
for sms in 8 12 16 24 32; do
echo "chain sms $sms"
AZN_NMS_CHAIN_SMS=$sms timeout 300 python tools/microbench.py --only nms --sizes 8000,20000 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['N'], d['thresh'], round(d['ms_best'],4), round(d['boxes_per_s']/1e6,1))
"
done
