cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --only c2 --only-step > gpurun_out/r02_bench_n2_v3.json 2> gpurun_out/r02_bench_n2_v3.err; echo "bench overlap rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --no-overlap --only c2 --only-step > gpurun_out/r02_bench_n2_v3_fused.json 2> gpurun_out/r02_bench_n2_v3_fused.err; echo "bench fused rc=$?"
CUDA_VISIBLE_DEVICES=1 python bench.py --only c2 --only-step > gpurun_out/r02_bench_n1_gpu1.json 2> gpurun_out/r02_bench_n1_gpu1.err; echo "gpu1 alone rc=$?"
CUDA_VISIBLE_DEVICES=0 python bench.py --only c2 --only-step > gpurun_out/r02_bench_n1_gpu0.json 2> gpurun_out/r02_bench_n1_gpu0.err; echo "gpu0 alone rc=$?"
