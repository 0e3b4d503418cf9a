#!/bin/bash
cd $GRAFT_REPO_ROOT
show() { python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k,v in d.get('configs',{}).items():
    print(k, 'ms', round(v.get('ms_per_step',0),4), 'value', round(v.get('value',0)), 'b1', (v.get('batch1') or {}).get('latency_us'), 'frac_dense', round((v.get('roofline') or {}).get('frac_dense',0),3), 'parity', (v.get('parity') or {}).get('ok'))
"; }
python bench.py --only c3,c4 --no-cpu-baseline 2>/dev/null | show
python bench.py --only c1 --no-cpu-baseline 2>/dev/null | show
python bench.py --only c5 --c5-global-batch 64 --no-cpu-baseline 2>/dev/null | show
