#!/bin/bash
# usage: bash scripts/gpurun_r02_final_multi_short.sh N   — headline + c5 at N GPUs with the final kernels (step times only)
cd $GRAFT_REPO_ROOT
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --only c2,c5 --only-step > gpurun_out/r02_bench_n${N}_final_steps.json 2> gpurun_out/r02_bench_n${N}_final_steps.err; echo "bench N=$N rc=$?"
tail -2 gpurun_out/r02_bench_n${N}_final_steps.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_bench_n${N}_final_steps.json').read().strip().splitlines()[-1])
print('headline', round(d['value']), d['ms_per_step'], d['n_gpus'], d['config'].get('exchange_status'))
c=d['configs']['c5']; print('c5', round(c['value']), c['ms_per_step'], c['per_gpu_batch'], c.get('exchange'))
"
