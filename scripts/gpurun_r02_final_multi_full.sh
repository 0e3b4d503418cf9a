#!/bin/bash
# usage: bash scripts/gpurun_r02_final_multi_full.sh N   — the driver-style full bench line at N GPUs
cd $GRAFT_REPO_ROOT
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}_final.json 2> gpurun_out/r02_bench_n${N}_final.err; echo "bench N=$N rc=$?"
tail -2 gpurun_out/r02_bench_n${N}_final.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n${N}_final.json').read().strip().splitlines()[-1])
print('headline', round(d['value']), d['ms_per_step'], d['n_gpus'], 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3))
c=d['configs']['c5']; print('c5', round(c['value']), c['ms_per_step'], c.get('parity'))
print(sorted(d.keys()))
"
