#!/bin/bash
# Last check of the round at the final HEAD: full GPU suite + smoke + a short bench line.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'], 4), round(d['e2e']['value']), d['roofline']['frac'], d['gpu_launches'])"
