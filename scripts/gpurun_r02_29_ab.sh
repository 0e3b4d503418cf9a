#!/bin/bash
# A/B of an environment switch on c1 / c5 / c5 at 64 images: bash scripts/gpurun_r02_29_ab.sh VAR
cd $GRAFT_REPO_ROOT
V=$1
show() { python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k,v in d.get('configs',{}).items():
    print('   ', k, 'ms', round(v.get('ms_per_step',0),4), 'b1', (v.get('batch1') or {}).get('latency_us'))
"; }
for rep in 1 2; do for val in 0 1; do
  echo "== $V=$val (pass $rep)"
  env $V=$val python bench.py --only c1 --no-cpu-baseline 2>/dev/null | show
  env $V=$val python bench.py --only c5 --no-cpu-baseline 2>/dev/null | show
  env $V=$val python bench.py --only c5 --c5-global-batch 64 --no-cpu-baseline 2>/dev/null | show
done; done
