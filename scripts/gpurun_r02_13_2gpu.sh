cd $GRAFT_REPO_ROOT
# (a) two independent single-GPU processes at the same time
CUDA_VISIBLE_DEVICES=0 python bench.py --only c2 --only-step > gpurun_out/r02_ctl_gpu0.json 2> gpurun_out/r02_ctl_gpu0.err &
P0=$!
CUDA_VISIBLE_DEVICES=1 python bench.py --only c2 --only-step > gpurun_out/r02_ctl_gpu1.json 2> gpurun_out/r02_ctl_gpu1.err &
P1=$!
wait $P0; wait $P1; echo "concurrent independent done"
# (b) torchrun N=2 without any exchange
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --only c2 --only-step --exchange local > gpurun_out/r02_ctl_n2_local.json 2> gpurun_out/r02_ctl_n2_local.err; echo "local rc=$?"
# (c) nccl in graph for comparison
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --only c2 --only-step --exchange nccl > gpurun_out/r02_ctl_n2_nccl.json 2> gpurun_out/r02_ctl_n2_nccl.err; echo "nccl rc=$?"
