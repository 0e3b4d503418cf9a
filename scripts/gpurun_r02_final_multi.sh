# usage: bash scripts/gpurun_r02_final_multi.sh N
cd $GRAFT_REPO_ROOT
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r02_bench_n${N}_final.json 2> gpurun_out/r02_bench_n${N}_final.err; echo "bench N=$N rc=$?"
tail -3 gpurun_out/r02_bench_n${N}_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --only c2 --only-step --fused-exchange --graph-steps 1 > gpurun_out/r02_bench_n${N}_fused.json 2> gpurun_out/r02_bench_n${N}_fused.err; echo "fused N=$N rc=$?"
