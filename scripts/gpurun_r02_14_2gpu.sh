cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_exchange.py -m gpu -x -q > gpurun_out/r02_pytest_14_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_14_2gpu.log
tail -4 gpurun_out/r02_pytest_14_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-overlap --only c2 --only-step > gpurun_out/r02_bench_n2_v4_fused.json 2> gpurun_out/r02_bench_n2_v4_fused.err; echo "bench fused rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --only c2 --only-step > gpurun_out/r02_bench_n2_v4.json 2> gpurun_out/r02_bench_n2_v4.err; echo "bench overlap rc=$?"
