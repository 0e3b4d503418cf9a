cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_exchange.py tests/test_gpu_yolo_loss.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --only c2 --only-step > gpurun_out/r02_bench_n2_v5_side.json 2> gpurun_out/r02_bench_n2_v5_side.err; echo "side rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --only c2 --only-step --fused-exchange > gpurun_out/r02_bench_n2_v5_fused.json 2> gpurun_out/r02_bench_n2_v5_fused.err; echo "fused rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err; echo "full rc=$?"
