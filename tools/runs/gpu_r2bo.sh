#!/bin/bash
# ROI-pool regression hunt: the same microbench with three builds of roi_pool.cu linked into the library --
# HEAD, 5877f7a (pair table compiled in, no geometry pre-pass) and ee33f6f (before both).  The alternative libraries are
# built in the authoring container into aznet_b200/build/ab/ (git-ignored, travels with the snapshot).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp aznet_b200/libaznet_b200.so /tmp/lib_head.so
for v in head old mid head; do
  if [ $v = head ]; then cp /tmp/lib_head.so aznet_b200/libaznet_b200.so; else cp aznet_b200/build/ab/lib_$v.so aznet_b200/libaznet_b200.so; fi
  touch aznet_b200/libaznet_b200.so
  for hw in 38,63 30,50; do
    timeout 200 python tools/microbench.py --only roi_pool --hw $hw 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d['variant'] != 'nchw_f32': print('$v', d['variant'], d['map'], d['R'], round(d['ms_best'], 4), round(d['ms_mean'], 4), round(d['frac_of_measured_hbm'], 3))
" | tee -a gpurun_out/r2bo_pool_ab.txt
  done
done
cp /tmp/lib_head.so aznet_b200/libaznet_b200.so
