cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m | head -8
python -m pytest tests/test_gpu_exchange.py -m gpu -x -q > gpurun_out/r02_pytest_7_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_7_2gpu.log
tail -15 gpurun_out/r02_pytest_7_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --repeats 9 --config-repeats 5 --verbose > gpurun_out/r02_bench_n2_v1.json 2> gpurun_out/r02_bench_n2_v1.err; echo "bench peer rc=$?"
tail -5 gpurun_out/r02_bench_n2_v1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --repeats 9 --config-repeats 5 --exchange nccl --only c2 > gpurun_out/r02_bench_n2_v1_nccl.json 2> gpurun_out/r02_bench_n2_v1_nccl.err; echo "bench nccl rc=$?"
tail -3 gpurun_out/r02_bench_n2_v1_nccl.err
